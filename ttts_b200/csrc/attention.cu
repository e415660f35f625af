// Causal flash attention, head_dim 64, forward and backward, for the UnifiedVoice GPT block.
// Replaces HF GPT2Attention (HF: modeling_gpt2.py:144-226; eager definition 54-72; sdpa path):
//   softmax(Q K^T * 64^-0.5 + causal) (+ attn dropout 0.1) @ V  on (B,H,T,64), T = 644 / 1156.
//
// Layout: qkv is the c_attn output [B*T, 3d] bf16 (q | k | v on the last dim, head h = columns h*64..h*64+63
// of each third, HF: modeling_gpt2.py:185-198), read in place with strides -- no head transposes; the output is
// written straight into the [B*T, d] buffer that c_proj consumes.
//
// Round-1 implementation: flash-style online softmax on the legacy warp-level tensor path (mma.sync.m16n8k16
// bf16, ldmatrix, cp.async double buffering), one CTA per (64-query block, head).  Backward is split into a
// dK/dV kernel (CTA owns a key block) and a dQ kernel (CTA owns a query block), so no atomics and bitwise
// run-to-run reproducible.  A tcgen05/TMEM version is the round-2 item.
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

constexpr int HD = 64;        // head dim
constexpr int BQ = 64;        // queries per CTA
constexpr int BKV = 64;       // keys per inner block
constexpr int ATT_THREADS = 128;
constexpr float kLog2e = 1.4426950408889634f;

// 64x64 bf16 tile, 128B rows, 16B chunks XOR-swizzled by (row & 7)
TTTS_DEVICE uint32_t tile_addr(uint32_t base, int row, int chunk) { return base + row * 128 + (((chunk ^ row) & 7) << 4); }

// async-load a [64 x 64] bf16 tile: rows row0.. of a matrix with row stride ld (elements); rows >= nrows zero-filled
TTTS_DEVICE void load_tile_async(uint32_t smem_base, const bf16* __restrict__ g, int ld, int row0, int nrows, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * ATT_THREADS;
        const int r = idx >> 3, c = idx & 7;
        const bool ok = (row0 + r) < nrows;
        const bf16* src = g + (size_t)(ok ? (row0 + r) : 0) * ld + c * 8;
        uint32_t dst = tile_addr(smem_base, r, c);
        uint32_t sz = ok ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
}

// A-operand fragments (16 rows x 64 k) for this warp's rows [wrow0, wrow0+16) of a swizzled tile
TTTS_DEVICE void load_a_frags(uint32_t base, int wrow0, int lane, uint32_t (&a)[4][4]) {
    const int m = lane >> 3, r = lane & 7;
    const int row = wrow0 + r + (m & 1) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(a[kk][0], a[kk][1], a[kk][2], a[kk][3], tile_addr(base, row, kk * 2 + (m >> 1)));
}
// B fragments from a tile stored [n][k] (row = n index, contiguous k): n-tiles (j, j+1), k-step kk
TTTS_DEVICE void load_b_nk(uint32_t base, int j, int kk, int lane, uint32_t (&b0)[2], uint32_t (&b1)[2]) {
    const int m = lane >> 3, r = lane & 7;
    ldmatrix_x4(b0[0], b0[1], b1[0], b1[1], tile_addr(base, (j + (m >> 1)) * 8 + r, kk * 2 + (m & 1)));
}
// B fragments from a tile stored [k][n] (row = k index, contiguous n): n-tiles (j, j+1), k-step kk
TTTS_DEVICE void load_b_kn(uint32_t base, int j, int kk, int lane, uint32_t (&b0)[2], uint32_t (&b1)[2]) {
    const int m = lane >> 3, r = lane & 7;
    ldmatrix_x4_trans(b0[0], b0[1], b1[0], b1[1], tile_addr(base, kk * 16 + (m & 1) * 8 + r, j + (m >> 1)));
}

// acc[8][4] (16 x 64) += A(16x64, regs) * B^T where B tile is [n=64][k=64]
TTTS_DEVICE void mma_a_bnk(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t bbase, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            uint32_t b0[2], b1[2];
            load_b_nk(bbase, j, kk, lane, b0, b1);
            mma_bf16_16816(acc[j], a[kk], b0);
            mma_bf16_16816(acc[j + 1], a[kk], b1);
        }
    }
}
// acc[8][4] (16 x 64) += P(16x64 in C-fragment layout, converted to bf16 A frags) * B where B tile is [k=64][n=64]
TTTS_DEVICE void mma_p_bkn(float (&acc)[8][4], const uint32_t (&pa)[4][4], uint32_t bbase, int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            uint32_t b0[2], b1[2];
            load_b_kn(bbase, j, kk, lane, b0, b1);
            mma_bf16_16816(acc[j], pa[kk], b0);
            mma_bf16_16816(acc[j + 1], pa[kk], b1);
        }
    }
}
// C-fragment (16x64 fp32) -> A fragments (bf16) for a following MMA whose k index is this tile's column index
TTTS_DEVICE void c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        a[kk][0] = pack_bf16(c[2 * kk][0], c[2 * kk][1]);
        a[kk][1] = pack_bf16(c[2 * kk][2], c[2 * kk][3]);
        a[kk][2] = pack_bf16(c[2 * kk + 1][0], c[2 * kk + 1][1]);
        a[kk][3] = pack_bf16(c[2 * kk + 1][2], c[2 * kk + 1][3]);
    }
}

// dropout on a pair of adjacent probabilities (row i, cols j, j+1; j even) of head-batch bh: same keep function as the tcgen05
// kernels (common.cuh attn_drop_*: row key from (seed, bh*T + i), one 3-round multiply-fold per group of 4 keys)
TTTS_DEVICE void drop_pair(const DropCfg& drop, uint64_t bh, int T, int i, int j, float& p0, float& p1) {
    uint32_t w0, w1;
    attn_drop_words(attn_drop_row(drop.seed, bh * (uint64_t)T + (uint64_t)i), (uint32_t)j >> 2, w0, w1);
    const uint32_t t32 = drop.thresh16 << 16;
    p0 = attn_drop_keep(w0, w1, j & 3, t32) ? p0 * drop.scale : 0.f;
    p1 = attn_drop_keep(w0, w1, (j & 3) + 1, t32) ? p1 * drop.scale : 0.f;      // j even -> j+1 is in the same group of 4
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse_out,
                                                               int T, int H, float scale, DropCfg drop) {
    __shared__ __align__(128) uint8_t smem[5 * 8192];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qb = gridDim.x - 1 - blockIdx.x;   // heavy (late) query blocks first
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int d = H * HD, ld = 3 * d;
    const bf16* qbase = qkv + (size_t)b * T * ld + h * HD;
    const bf16* kbase = qbase + d;
    const bf16* vbase = qbase + 2 * d;
    const uint32_t sQ = smem_u32(smem), sK = sQ + 8192, sV = sK + 2 * 8192;
    const int q0 = qb * BQ;
    const int nkv = qb + 1;   // causal: key blocks 0..qb

    load_tile_async(sQ, qbase, ld, q0, T, tid);
    load_tile_async(sK, kbase, ld, 0, T, tid);
    load_tile_async(sV, vbase, ld, 0, T, tid);
    cp_async_commit();

    uint32_t qa[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int g = lane >> 2, t4 = lane & 3;
    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
    const float sl2 = scale * kLog2e;

    for (int n = 0; n < nkv; ++n) {
        const int st = n & 1;
        if (n + 1 < nkv) {
            load_tile_async(sK + (st ^ 1) * 8192, kbase, ld, (n + 1) * BKV, T, tid);
            load_tile_async(sV + (st ^ 1) * 8192, vbase, ld, (n + 1) * BKV, T, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (n == 0) load_a_frags(sQ, warp * 16, lane, qa);

        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
        mma_a_bnk(s, qa, sK + st * 8192, lane);

        // mask (diagonal block and the key tail)
        const int k0 = n * BKV;
        if (n == nkv - 1 || k0 + BKV > T) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = k0 + j * 8 + t4 * 2;
                if (c > row_a || c >= T) s[j][0] = -INFINITY;
                if (c + 1 > row_a || c + 1 >= T) s[j][1] = -INFINITY;
                if (c > row_b || c >= T) s[j][2] = -INFINITY;
                if (c + 1 > row_b || c + 1 >= T) s[j][3] = -INFINITY;
            }
        }
        // online softmax (rows g and g+8 of this warp's 16)
        float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            msc[r] = (mx[r] == -INFINITY) ? 0.f : mx[r] * sl2;
            corr[r] = exp2f(m_run[r] * sl2 - msc[r]);      // m_run = -inf -> 0
            m_run[r] = mx[r];
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(s[j][0] * sl2 - msc[0]);
            s[j][1] = exp2f(s[j][1] * sl2 - msc[0]);
            s[j][2] = exp2f(s[j][2] * sl2 - msc[1]);
            s[j][3] = exp2f(s[j][3] * sl2 - msc[1]);
            rs[0] += s[j][0] + s[j][1];
            rs[1] += s[j][2] + s[j][3];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr[0]; o[j][1] *= corr[0];
            o[j][2] *= corr[1]; o[j][3] *= corr[1];
        }
        if (drop.thresh16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = k0 + j * 8 + t4 * 2;
                drop_pair(drop, (uint64_t)bh, T, row_a, c, s[j][0], s[j][1]);
                drop_pair(drop, (uint64_t)bh, T, row_b, c, s[j][2], s[j][3]);
            }
        }
        uint32_t pa[4][4];
        c_to_a(s, pa);
        mma_p_bkn(o, pa, sV + st * 8192, lane);
        __syncthreads();
    }

    // finalize: O /= l ; lse = m*scale + ln(l)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    if (t4 == 0) {
        if (row_a < T) lse_out[(size_t)bh * T + row_a] = m_run[0] * scale + logf(l_run[0]);
        if (row_b < T) lse_out[(size_t)bh * T + row_b] = m_run[1] * scale + logf(l_run[1]);
    }
    // stage through smem (Q tile is dead) for 16B coalesced stores
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ra = warp * 16 + g, rb = ra + 8;
        const uint32_t a0 = tile_addr(sQ, ra, j) + t4 * 4;
        const uint32_t a1 = tile_addr(sQ, rb, j) + t4 * 4;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(pack_bf16(o[j][0] * inv0, o[j][1] * inv0)) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(pack_bf16(o[j][2] * inv1, o[j][3] * inv1)) : "memory");
    }
    __syncthreads();
    bf16* obase = out + (size_t)b * T * d + h * HD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * ATT_THREADS;
        const int r = idx >> 3, c = idx & 7;
        if (q0 + r < T) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_addr(sQ, r, c)));
            *reinterpret_cast<uint4*>(obase + (size_t)(q0 + r) * d + c * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward pre-pass: delta[bh, i] = sum_d dO[i,d] * O[i,d]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                                         int B, int T, int H) {
    pdl_launch_dependents();
    pdl_wait();
    // 8 lanes per (token, head): one 16-byte load of O and of dO each (8 dims), 3 shuffle steps; 2 pairs per thread for loads in flight
    const size_t total = (size_t)B * T * H;
    const size_t pair0 = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const size_t half = (total + 1) / 2;
    float s[2] = {0.f, 0.f};
    size_t pr[2] = {pair0, pair0 + half};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (pair0 >= half || pr[u] >= total) continue;
        const size_t off = pr[u] * HD + sub * 8;          // pairs are (token, head) row-major: token * (H*64) + h*64
        const uint4 a = *reinterpret_cast<const uint4*>(o + off);
        const uint4 g = *reinterpret_cast<const uint4*>(dout + off);
        s[u] = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) + bf16_hi(a.y) * bf16_hi(g.y) +
               bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) + bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        float v = s[u];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sub == 0 && pair0 < half && pr[u] < total) {
            const size_t tok = pr[u] / H; const int h = (int)(pr[u] - tok * H);
            const size_t b = tok / T; const int t = (int)(tok - b * T);
            delta[(b * H + h) * T + t] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, dK/dV: CTA owns key block nb; warp owns 16 keys; loops over query blocks m >= nb
//   S^T = K Q^T ; P^T = exp(S^T*scale - lse[q]) ; dV += Pdrop^T dO ; dP^T = V dO^T ; dS^T = P^T (dP^T - delta[q]) scale ; dK += dS^T Q
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkdv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                    const float* __restrict__ lse, const float* __restrict__ delta,
                                                                    bf16* __restrict__ dqkv, int T, int H, float scale, DropCfg drop) {
    extern __shared__ __align__(128) uint8_t dsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = blockIdx.x;                    // key block (early key blocks are the heavy ones: natural order)
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int d = H * HD, ld = 3 * d;
    const bf16* qbase = qkv + (size_t)b * T * ld + h * HD;
    const bf16* kbase = qbase + d;
    const bf16* vbase = qbase + 2 * d;
    const bf16* dobase = dout + (size_t)b * T * d + h * HD;
    const uint32_t sK = smem_u32(dsm), sV = sK + 8192, sQ = sV + 8192, sDO = sQ + 2 * 8192;
    float* sLse = reinterpret_cast<float*>(dsm + 6 * 8192);       // [2][64]
    float* sDel = sLse + 128;                                     // [2][64]
    const int k0 = nb * BKV;
    const int nq = (T + BQ - 1) / BQ;

    load_tile_async(sK, kbase, ld, k0, T, tid);
    load_tile_async(sV, vbase, ld, k0, T, tid);
    load_tile_async(sQ, qbase, ld, nb * BQ, T, tid);
    load_tile_async(sDO, dobase, d, nb * BQ, T, tid);
    cp_async_commit();
    if (tid < 64) {
        const int q = nb * BQ + tid;
        sLse[tid] = q < T ? lse[(size_t)bh * T + q] : 0.f;
        sDel[tid] = q < T ? delta[(size_t)bh * T + q] : 0.f;
    }

    uint32_t ka[4][4], va[4][4];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f; dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f; }
    const int g = lane >> 2, t4 = lane & 3;
    const int key_a = k0 + warp * 16 + g, key_b = key_a + 8;
    const float sl2 = scale * kLog2e;

    for (int m = nb; m < nq; ++m) {
        const int st = (m - nb) & 1;
        if (m + 1 < nq) {
            load_tile_async(sQ + (st ^ 1) * 8192, qbase, ld, (m + 1) * BQ, T, tid);
            load_tile_async(sDO + (st ^ 1) * 8192, dobase, d, (m + 1) * BQ, T, tid);
            cp_async_commit();
            if (tid < 64) {
                const int q = (m + 1) * BQ + tid;
                sLse[(st ^ 1) * 64 + tid] = q < T ? lse[(size_t)bh * T + q] : 0.f;
                sDel[(st ^ 1) * 64 + tid] = q < T ? delta[(size_t)bh * T + q] : 0.f;
            }
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (m == nb) { load_a_frags(sK, warp * 16, lane, ka); load_a_frags(sV, warp * 16, lane, va); }
        const uint32_t cQ = sQ + st * 8192, cDO = sDO + st * 8192;
        const float* cl = sLse + st * 64;
        const float* cd = sDel + st * 64;
        const int q0 = m * BQ;

        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
        mma_a_bnk(s, ka, cQ, lane);            // S^T[key, query]
        float dp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
        mma_a_bnk(dp, va, cDO, lane);          // dP^T[key, query]

        uint32_t pa[4][4];                     // dropped probabilities as A fragments (for dV)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int qc = q0 + j * 8 + t4 * 2;     // query index of columns (qc, qc+1)
            const float l0 = cl[j * 8 + t4 * 2] * kLog2e, l1 = cl[j * 8 + t4 * 2 + 1] * kLog2e;
            const float d0 = cd[j * 8 + t4 * 2], d1 = cd[j * 8 + t4 * 2 + 1];
            float p[4];
            p[0] = (key_a <= qc && qc < T && key_a < T) ? exp2f(s[j][0] * sl2 - l0) : 0.f;
            p[1] = (key_a <= qc + 1 && qc + 1 < T && key_a < T) ? exp2f(s[j][1] * sl2 - l1) : 0.f;
            p[2] = (key_b <= qc && qc < T && key_b < T) ? exp2f(s[j][2] * sl2 - l0) : 0.f;
            p[3] = (key_b <= qc + 1 && qc + 1 < T && key_b < T) ? exp2f(s[j][3] * sl2 - l1) : 0.f;
            float q4[4] = {p[0], p[1], p[2], p[3]};
            float g4[4] = {dp[j][0], dp[j][1], dp[j][2], dp[j][3]};
            if (drop.thresh16) {
                // mask index is (query row, key col): element (i = qc(+1), j = key)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qi = qc + (e & 1), kj = (e < 2) ? key_a : key_b;
                    const float mk = attn_drop_keep1(drop.seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi, kj, drop.thresh16) ? drop.scale : 0.f;
                    q4[e] *= mk;
                    g4[e] *= mk;
                }
            }
            pa[j >> 1][(j & 1) * 2] = pack_bf16(q4[0], q4[1]);
            pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(q4[2], q4[3]);
            s[j][0] = p[0] * (g4[0] - d0) * scale;
            s[j][1] = p[1] * (g4[1] - d1) * scale;
            s[j][2] = p[2] * (g4[2] - d0) * scale;
            s[j][3] = p[3] * (g4[3] - d1) * scale;
        }
        mma_p_bkn(dv, pa, cDO, lane);          // dV += Pdrop^T dO   (B = dO [query][dim] = [k][n])
        c_to_a(s, pa);
        mma_p_bkn(dk, pa, cQ, lane);           // dK += dS^T Q
        __syncthreads();
    }

    // store dK, dV (bf16) via smem staging (K/V tiles are dead)
    bf16* dkbase = dqkv + (size_t)b * T * ld + d + h * HD;
    bf16* dvbase = dkbase + d;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ra = warp * 16 + g, rb = ra + 8;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(sK, ra, j) + t4 * 4), "r"(pack_bf16(dk[j][0], dk[j][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(sK, rb, j) + t4 * 4), "r"(pack_bf16(dk[j][2], dk[j][3])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(sV, ra, j) + t4 * 4), "r"(pack_bf16(dv[j][0], dv[j][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(sV, rb, j) + t4 * 4), "r"(pack_bf16(dv[j][2], dv[j][3])) : "memory");
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * ATT_THREADS;
        const int r = idx >> 3, c = idx & 7;
        if (k0 + r < T) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_addr(sK, r, c)));
            *reinterpret_cast<uint4*>(dkbase + (size_t)(k0 + r) * ld + c * 8) = v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_addr(sV, r, c)));
            *reinterpret_cast<uint4*>(dvbase + (size_t)(k0 + r) * ld + c * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, dQ: CTA owns query block; warp owns 16 queries; loops over key blocks n <= qb
//   S = Q K^T ; P = exp(S*scale - lse) ; dP = dO V^T (masked by dropout) ; dS = P (dP - delta) scale ; dQ += dS K
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                  const float* __restrict__ lse, const float* __restrict__ delta,
                                                                  bf16* __restrict__ dqkv, int T, int H, float scale, DropCfg drop) {
    extern __shared__ __align__(128) uint8_t dsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qb = gridDim.x - 1 - blockIdx.x;
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int d = H * HD, ld = 3 * d;
    const bf16* qbase = qkv + (size_t)b * T * ld + h * HD;
    const bf16* kbase = qbase + d;
    const bf16* vbase = qbase + 2 * d;
    const bf16* dobase = dout + (size_t)b * T * d + h * HD;
    const uint32_t sQ = smem_u32(dsm), sDO = sQ + 8192, sK = sDO + 8192, sV = sK + 2 * 8192;
    const int q0 = qb * BQ;
    const int nkv = qb + 1;

    load_tile_async(sQ, qbase, ld, q0, T, tid);
    load_tile_async(sDO, dobase, d, q0, T, tid);
    load_tile_async(sK, kbase, ld, 0, T, tid);
    load_tile_async(sV, vbase, ld, 0, T, tid);
    cp_async_commit();

    const int g = lane >> 2, t4 = lane & 3;
    const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
    const float lse_a = (row_a < T ? lse[(size_t)bh * T + row_a] : 0.f) * kLog2e;
    const float lse_b = (row_b < T ? lse[(size_t)bh * T + row_b] : 0.f) * kLog2e;
    const float del_a = row_a < T ? delta[(size_t)bh * T + row_a] : 0.f;
    const float del_b = row_b < T ? delta[(size_t)bh * T + row_b] : 0.f;
    const float sl2 = scale * kLog2e;

    uint32_t qa[4][4], doa[4][4];
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }

    for (int n = 0; n < nkv; ++n) {
        const int st = n & 1;
        if (n + 1 < nkv) {
            load_tile_async(sK + (st ^ 1) * 8192, kbase, ld, (n + 1) * BKV, T, tid);
            load_tile_async(sV + (st ^ 1) * 8192, vbase, ld, (n + 1) * BKV, T, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (n == 0) { load_a_frags(sQ, warp * 16, lane, qa); load_a_frags(sDO, warp * 16, lane, doa); }
        const uint32_t cK = sK + st * 8192, cV = sV + st * 8192;
        const int k0 = n * BKV;

        float s[8][4], dp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
        mma_a_bnk(s, qa, cK, lane);            // S[query, key]
        mma_a_bnk(dp, doa, cV, lane);          // dP[query, key] = dO V^T
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = k0 + j * 8 + t4 * 2;
            float p[4];
            p[0] = (c <= row_a && c < T && row_a < T) ? exp2f(s[j][0] * sl2 - lse_a) : 0.f;
            p[1] = (c + 1 <= row_a && c + 1 < T && row_a < T) ? exp2f(s[j][1] * sl2 - lse_a) : 0.f;
            p[2] = (c <= row_b && c < T && row_b < T) ? exp2f(s[j][2] * sl2 - lse_b) : 0.f;
            p[3] = (c + 1 <= row_b && c + 1 < T && row_b < T) ? exp2f(s[j][3] * sl2 - lse_b) : 0.f;
            float g4[4] = {dp[j][0], dp[j][1], dp[j][2], dp[j][3]};
            if (drop.thresh16) {
                float one0 = 1.f, one1 = 1.f, one2 = 1.f, one3 = 1.f;
                drop_pair(drop, (uint64_t)bh, T, row_a, c, one0, one1);
                drop_pair(drop, (uint64_t)bh, T, row_b, c, one2, one3);
                g4[0] *= one0; g4[1] *= one1; g4[2] *= one2; g4[3] *= one3;
            }
            s[j][0] = p[0] * (g4[0] - del_a) * scale;
            s[j][1] = p[1] * (g4[1] - del_a) * scale;
            s[j][2] = p[2] * (g4[2] - del_b) * scale;
            s[j][3] = p[3] * (g4[3] - del_b) * scale;
        }
        uint32_t pa[4][4];
        c_to_a(s, pa);
        mma_p_bkn(dq, pa, cK, lane);           // dQ += dS K   (B = K [key][dim] = [k][n])
        __syncthreads();
    }

    bf16* dqbase = dqkv + (size_t)b * T * ld + h * HD;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ra = warp * 16 + g, rb = ra + 8;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(sQ, ra, j) + t4 * 4), "r"(pack_bf16(dq[j][0], dq[j][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_addr(sQ, rb, j) + t4 * 4), "r"(pack_bf16(dq[j][2], dq[j][3])) : "memory");
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * ATT_THREADS;
        const int r = idx >> 3, c = idx & 7;
        if (q0 + r < T) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_addr(sQ, r, c)));
            *reinterpret_cast<uint4*>(dqbase + (size_t)(q0 + r) * ld + c * 8) = v;
        }
    }
}

// attention_tc.cu (tcgen05 path)
bool attn_use_tc();
int attn_fwd_tc(const bf16* qkv, bf16* o, float* lse, int B, int T, int H, DropCfg drop, cudaStream_t st);
int attn_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int B, int T, int H, DropCfg drop,
                cudaStream_t st);

int attn_delta(const bf16* o, const bf16* dout, float* delta, int B, int T, int H, cudaStream_t st) {
    const size_t half = ((size_t)B * T * H + 1) / 2;                       // each thread group of 8 lanes handles pairs p and p + half
    TTTS_CUDA(launch_pdl(attn_delta_kernel, dim3((unsigned)((half * 8 + 255) / 256)), dim3(256), 0, st, o, dout, delta, B, T, H));
    TTTS_LAUNCH_CHECK("attn_delta");
    return TTTS_OK;
}

int attn_fwd(const bf16* qkv, bf16* o, float* lse, int B, int T, int H, DropCfg drop, cudaStream_t st) {
    TTTS_CHECK_ARG(B > 0 && T > 0 && H > 0, "attn: bad shape");
    TTTS_CHECK_ARG((size_t)B * H <= 65535, "attn: B*H too large for grid.y");
    if (attn_use_tc()) return attn_fwd_tc(qkv, o, lse, B, T, H, drop, st);
    dim3 grid((T + BQ - 1) / BQ, B * H);
    attn_fwd_kernel<<<grid, ATT_THREADS, 0, st>>>(qkv, o, lse, T, H, 0.125f, drop);
    TTTS_LAUNCH_CHECK("attn_fwd");
    return TTTS_OK;
}

int attn_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int B, int T, int H, DropCfg drop,
             cudaStream_t st) {
    TTTS_CHECK_ARG(B > 0 && T > 0 && H > 0, "attn: bad shape");
    TTTS_CHECK_ARG((size_t)B * H <= 65535, "attn: B*H too large for grid.y");
    if (attn_use_tc()) return attn_bwd_tc(qkv, o, dout, lse, delta, dqkv, B, T, H, drop, st);
    { int rc = attn_delta(o, dout, delta, B, T, H, st); if (rc) return rc; }
    dim3 grid((T + BQ - 1) / BQ, B * H);
    const int smem_kv = 6 * 8192 + 4 * 64 * 4;
    const int smem_q = 6 * 8192;
    static bool attr = false;
    if (!attr) {
        TTTS_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv));
        TTTS_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
        attr = true;
    }
    attn_bwd_dkdv_kernel<<<grid, ATT_THREADS, smem_kv, st>>>(qkv, dout, lse, delta, dqkv, T, H, 0.125f, drop);
    TTTS_LAUNCH_CHECK("attn_bwd_dkdv");
    attn_bwd_dq_kernel<<<grid, ATT_THREADS, smem_q, st>>>(qkv, dout, lse, delta, dqkv, T, H, 0.125f, drop);
    TTTS_LAUNCH_CHECK("attn_bwd_dq");
    return TTTS_OK;
}

// keep mask of the attention-probability dropout, [BH, T, T] bytes (tests: lets torch reproduce the dropped attention exactly)
__global__ void attn_dropout_mask_kernel(uint8_t* __restrict__ mask, int T, DropCfg drop) {
    const int qi = blockIdx.x, bh = blockIdx.y;
    for (int kj = threadIdx.x; kj < T; kj += blockDim.x)
        mask[((size_t)bh * T + qi) * T + kj] = (!drop.thresh16 || attn_drop_keep1(drop.seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi, kj, drop.thresh16)) ? 1 : 0;
}
int attn_dropout_mask(uint8_t* mask, int BH, int T, DropCfg drop, cudaStream_t st) {
    TTTS_CHECK_ARG(mask != nullptr && BH > 0 && T > 0, "attn_dropout_mask: bad arguments");
    attn_dropout_mask_kernel<<<dim3(T, BH), 128, 0, st>>>(mask, T, drop);
    TTTS_LAUNCH_CHECK("attn_dropout_mask");
    return TTTS_OK;
}

// keep mask of an element-wise dropout site (embedding, attention-output, MLP-output), [rows, cols] bytes: the decision of element
// (row, c) is field c & 3 of dropout_bits4(seed, (row * cols + c) >> 2) -- the indexing every kernel of those sites uses
__global__ void elem_dropout_mask_kernel(uint8_t* __restrict__ mask, int cols, DropCfg drop) {
    const size_t row = blockIdx.x;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        const uint64_t bits = dropout_bits4(drop.seed, (row * (uint64_t)cols + (uint64_t)c) >> 2);
        mask[row * cols + c] = (!drop.thresh16 || dropout_keep(bits, c & 3, drop.thresh16)) ? 1 : 0;
    }
}
int elem_dropout_mask(uint8_t* mask, int rows, int cols, DropCfg drop, cudaStream_t st) {
    TTTS_CHECK_ARG(mask != nullptr && rows > 0 && cols > 0 && cols % 4 == 0, "elem_dropout_mask: bad arguments");
    elem_dropout_mask_kernel<<<rows, 256, 0, st>>>(mask, cols, drop);
    TTTS_LAUNCH_CHECK("elem_dropout_mask");
    return TTTS_OK;
}

}  // namespace ttts
