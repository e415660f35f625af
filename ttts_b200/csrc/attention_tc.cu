// Causal flash attention on the 5th-gen tensor cores (tcgen05.mma + TMEM + TMA), head_dim 64, forward and backward.
// Replaces HF GPT2Attention (HF: modeling_gpt2.py:144-226) like attention.cu, which stays as the legacy mma.sync path
// (TTTS_ATTN_LEGACY=1) and as a cross-check in the tests.
//
// Every operand tile is [128 rows x 64 bf16] = 128 rows of 128 B in the 128B-swizzle layout, loaded by ONE TMA box straight
// out of the packed c_attn output [B*T, 3d] (or the [B*T, d] dO buffer).  The same bytes serve as a K-major operand (rows =
// M/N index, the 64 head dims = K) and as an MN-major operand (rows = K index, the 64 head dims = M/N), so no transposes:
//
//   forward   S  = Q K^T        A=Q  (K-maj)  B=K (K-maj)   128x128x64  -> TMEM
//             P  = softmax tile (4 warps, one query row per thread = one TMEM lane: no shuffles), bf16 -> smem
//             O += P V          A=P  (K-maj)  B=V (MN-maj)  128x64x128  -> TMEM scratch, folded into registers
//   backward (CTA owns a key block, loops over query blocks):
//             S  = Q K^T , dP = dO V^T                       128x128x64 each -> TMEM
//             P = exp(S*scale - lse), dS = P (dP - delta) scale   (thread = query row) -> bf16 smem, both
//             dV += P^T dO      A=P  (MN-maj) B=dO (MN-maj) 128x64x128
//             dK += dS^T Q      A=dS (MN-maj) B=Q  (MN-maj) 128x64x128
//             dQ  = dS K        A=dS (K-maj)  B=K  (MN-maj) 128x64x128 -> TMEM -> red.global.add.f32 into an fp32 dQ buffer
//
// Warp roles (320 threads): warp 0 = TMA loader, warp 1 = MMA issuer + TMEM alloc, warps 2-9 = softmax / gradient math: two warps
// per TMEM lane quadrant, each owning half of the tile's columns (64 keys, 32 of the 64 output dims), so every SM sub-partition has
// two math warps to hide latency (the first version with one warp per sub-partition issued 0.19 instr/cycle, profiles/r1_notes.md).
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

constexpr int AT_BM = 128;            // queries per tile
constexpr int AT_BN = 128;            // keys per tile
constexpr int AT_TILE = 128 * 128;    // bytes of a [128 x 64] bf16 tile
constexpr int AT_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9: two warps per TMEM lane quadrant (each takes half of the columns)
constexpr float kLog2eF = 1.4426950408889634f;

TTTS_DEVICE void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
TTTS_DEVICE float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// K-major descriptor for k-step kk (16 elements of K) inside a [rows x 64] tile (or the 2-atom [rows x 128] P/dS buffers)
TTTS_DEVICE uint64_t desc_kmajor(uint32_t base, int kk) { return make_smem_desc_sw128(base + (kk >> 2) * AT_TILE + (kk & 3) * 32, 16, 1024); }
// MN-major descriptor for k-step kk (16 rows of K); MN extent 64 (single atom) or 128 (two atoms, AT_TILE apart)
TTTS_DEVICE uint64_t desc_mnmajor(uint32_t base, int kk) { return make_smem_desc_sw128(base + kk * 2048, AT_TILE, 1024); }

// write 8 bf16 (16 B) of row r, 16B-chunk c16 (0..15 over 128 columns) into a 2-atom K-major/MN-major swizzled buffer
TTTS_DEVICE void st_tile_chunk(uint32_t base, int r, int c16, uint4 v) {
    const uint32_t addr = base + (c16 >> 3) * AT_TILE + r * 128 + ((((c16 & 7) ^ r) & 7) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
struct FwdSmem {
    static constexpr int kKvStages = 3;
    static constexpr int oQ = 0;
    static constexpr int oKV = AT_TILE;                                 // [stages][K | V]
    static constexpr int oP = oKV + kKvStages * 2 * AT_TILE;            // [2][2 atoms]
    static constexpr int oBar = oP + 2 * 2 * AT_TILE;
    static constexpr int oXch = oBar + 256;                             // row-max / row-sum exchange between the two column halves: [2][2][128] floats
    static constexpr int kBytes = oXch + 3 * 2 * 128 * 4 + 1024;       // 2 max buffers + 1 sum buffer
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out, int T, int H, float scale,
                   DropCfg drop) {
    using S = FwdSmem;
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::oBar);
    uint64_t* q_full = bars;                       // 1
    uint64_t* kv_full = bars + 1;                  // [3]
    uint64_t* kv_empty = bars + 4;                 // [3]
    uint64_t* s_full = bars + 7;                   // [2]
    uint64_t* s_empty = bars + 9;                  // [2]
    uint64_t* p_full = bars + 11;                  // [2]
    uint64_t* p_empty = bars + 13;                 // [2]
    uint64_t* o_full = bars + 15;                  // [2]
    uint64_t* o_empty = bars + 17;                 // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = gridDim.x - 1 - blockIdx.x;            // heavy (late) query blocks first
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int d = H * 64;
    const int q0 = qb * AT_BM;
    const int nkv = min(qb + 1, (T + AT_BN - 1) / AT_BN);
    const int row_base = b * T;                            // row of token 0 of this sequence in the [B*T, .] matrices

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < S::kKvStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 256);
            mbar_init(&p_full[s], 256); mbar_init(&p_empty[s], 1);
            mbar_init(&o_full[s], 1); mbar_init(&o_empty[s], 256);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tS = tmem_base;              // [2] x 128 cols
    const uint32_t tO = tmem_base + 256;        // [2] x 64 cols

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA loader ----------------
            mbar_arrive_expect_tx(q_full, AT_TILE);
            tma_load_2d(smem + S::oQ, &tmQKV, q_full, h * 64, row_base + q0);
            int st = 0; uint32_t ph = 0;
            for (int j = 0; j < nkv; ++j) {
                mbar_wait(&kv_empty[st], ph ^ 1);
                uint8_t* sk = smem + S::oKV + st * 2 * AT_TILE;
                mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE);
                tma_load_2d(sk, &tmQKV, &kv_full[st], d + h * 64, row_base + j * AT_BN);
                tma_load_2d(sk + AT_TILE, &tmQKV, &kv_full[st], 2 * d + h * 64, row_base + j * AT_BN);
                if (++st == S::kKvStages) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
            const uint32_t sQ = smem_u32(smem + S::oQ);
            auto issue_s = [&](int j, int st) {
                const uint32_t sK = smem_u32(smem + S::oKV + st * 2 * AT_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tS + (j & 1) * 128, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);                                    // s_empty[0] is trivially free for j = 0
            int st = 0; uint32_t ph = 0;                      // stage / phase of block j
            for (int j = 0; j < nkv; ++j) {
                int st1 = st + 1; uint32_t ph1 = ph;
                if (st1 == S::kKvStages) { st1 = 0; ph1 ^= 1; }
                if (j + 1 < nkv) {
                    mbar_wait(&kv_full[st1], ph1);
                    mbar_wait(&s_empty[(j + 1) & 1], (((j + 1) >> 1) & 1) ^ 1);
                    tc_fence_after();
                    issue_s(j + 1, st1);
                }
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                mbar_wait(&o_empty[j & 1], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t sP = smem_u32(smem + S::oP + (j & 1) * 2 * AT_TILE);
                const uint32_t sV = smem_u32(smem + S::oKV + st * 2 * AT_TILE + AT_TILE);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_bf16(tO + (j & 1) * 64, desc_kmajor(sP, k), desc_mnmajor(sV, k), idesc_o, k > 0 ? 1u : 0u);
                umma_commit(&o_full[j & 1]);
                umma_commit(&p_empty[j & 1]);
                umma_commit(&kv_empty[st]);
                st = st1; ph = ph1;
            }
        }
        __syncwarp();
    } else {
        // ---------------- softmax warps: thread = (query row = TMEM lane, column half) ----------------
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                    // 0: keys 0-63 / out dims 0-31 ; 1: keys 64-127 / out dims 32-63
        const int r = quad * 32 + lane;
        const int qi = q0 + r;                               // query index within the sequence
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * kLog2eF;
        float* xch = reinterpret_cast<float*>(smem + S::oXch);               // [buf][half][row]
        float m_run = -INFINITY, l_run = 0.f;                // l_run: partial row sum over THIS thread's keys
        const AttnDropRow rk = attn_drop_row(drop.seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi);
        const uint32_t t32 = drop.thresh16 << 16;
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;

        for (int j = 0; j < nkv; ++j) {
            const int k0 = j * AT_BN;
            const bool need_mask = (j == qb) || (k0 + AT_BN > T);
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t ts = tS + (j & 1) * 128 + lane_off + half * 64;
            const int kc0 = k0 + half * 64;
            // pass 1: row max over this thread's 64 keys
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(ts + c * 32, v);
                tmem_ld_wait();
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kj = kc0 + c * 32 + i;
                        if (kj <= qi && kj < T) mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            xch[((j & 1) * 2 + half) * 128 + r] = mx;
            named_bar_sync(2, 256);
            mx = fmaxf(fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * 128 + r]), m_run);
            const float msc = (mx == -INFINITY) ? 0.f : mx * sl2;
            const float corr = ex2_fast(m_run * sl2 - msc);       // 0 when m_run = -inf
            // pass 2: probabilities -> bf16 P tile in smem
            mbar_wait(&p_empty[j & 1], ((j >> 1) & 1) ^ 1);
            const uint32_t sP = smem_u32(smem + S::oP + (j & 1) * 2 * AT_TILE);
            float rs = 0.f;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(ts + c * 32, v);
                tmem_ld_wait();
                float pv[32];
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kj = kc0 + c * 32 + i;
                        pv[i] = (kj <= qi && kj < T) ? ex2_fast(fmaf(__uint_as_float(v[i]), sl2, -msc)) : 0.f;
                        rs += pv[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { pv[i] = ex2_fast(fmaf(__uint_as_float(v[i]), sl2, -msc)); rs += pv[i]; }
                }
                if (drop.thresh16) {
                    // keep decisions of 4 consecutive keys per multiply-fold hash (common.cuh); the 1/(1-p) scale is applied once, to O
                    const uint32_t g0 = (uint32_t)(kc0 + c * 32) >> 2;
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        uint32_t w0, w1;
                        attn_drop_words(rk, g0 + i4, w0, w1);
                        pv[i4 * 4 + 0] = (w0 >= t32) ? pv[i4 * 4 + 0] : 0.f;
                        pv[i4 * 4 + 1] = ((w0 << 16) >= t32) ? pv[i4 * 4 + 1] : 0.f;
                        pv[i4 * 4 + 2] = (w1 >= t32) ? pv[i4 * 4 + 2] : 0.f;
                        pv[i4 * 4 + 3] = ((w1 << 16) >= t32) ? pv[i4 * 4 + 3] : 0.f;
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    st_tile_chunk(sP, r, (half * 2 + c) * 4 + g,
                                  make_uint4(pack_bf16(pv[8 * g], pv[8 * g + 1]), pack_bf16(pv[8 * g + 2], pv[8 * g + 3]),
                                             pack_bf16(pv[8 * g + 4], pv[8 * g + 5]), pack_bf16(pv[8 * g + 6], pv[8 * g + 7])));
            }
            tc_fence_before();
            mbar_arrive(&s_empty[j & 1]);
            fence_proxy_async();
            mbar_arrive(&p_full[j & 1]);
            l_run = l_run * corr + rs;
            m_run = mx;
            if (j >= 1) {        // fold in P_{j-1} V_{j-1}, which the tensor core finished while we did the softmax of block j
                mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(tO + ((j - 1) & 1) * 64 + lane_off + half * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v[i]);
                tc_fence_before();
                mbar_arrive(&o_empty[(j - 1) & 1]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= corr;
        }
        {
            const int jl = nkv - 1;
            mbar_wait(&o_full[jl & 1], (jl >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(tO + (jl & 1) * 64 + lane_off + half * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v[i]);
            tc_fence_before();
            mbar_arrive(&o_empty[jl & 1]);
        }
        // total row sum = the two halves' partial sums
        xch[(4 + half) * 128 + r] = l_run;
        named_bar_sync(2, 256);
        const float l_tot = l_run + xch[(4 + (half ^ 1)) * 128 + r];
        if (qi < T) {
            const float inv = l_tot > 0.f ? drop.scale / l_tot : 0.f;      // dropout's 1/(1-p) folded in here (scale = 1 when off)
            if (half == 0) lse_out[(size_t)bh * T + qi] = m_run * scale + logf(l_tot);
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(row_base + qi) * d + h * 64 + half * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                dst[g] = make_uint4(pack_bf16(o[8 * g] * inv, o[8 * g + 1] * inv), pack_bf16(o[8 * g + 2] * inv, o[8 * g + 3] * inv),
                                    pack_bf16(o[8 * g + 4] * inv, o[8 * g + 5] * inv), pack_bf16(o[8 * g + 6] * inv, o[8 * g + 7] * inv));
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------------------
// backward: CTA owns key block jb, loops over query blocks i >= jb
// ------------------------------------------------------------------------------------------------------------
struct BwdSmem {
    static constexpr int oK = 0;
    static constexpr int oV = AT_TILE;
    static constexpr int oQdO = 2 * AT_TILE;                  // [2 stages][Q | dO]
    static constexpr int oP = oQdO + 2 * 2 * AT_TILE;         // 2 atoms
    static constexpr int oDS = oP + 2 * AT_TILE;              // 2 atoms
    static constexpr int oBar = oDS + 2 * AT_TILE;
    static constexpr int kBytes = oBar + 256 + 1024;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const float* __restrict__ lse,
                   const float* __restrict__ delta, bf16* __restrict__ dqkv, float* __restrict__ dq_acc, int T, int H, float scale, DropCfg drop) {
    using S = BwdSmem;
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::oBar);
    uint64_t* kv_full = bars;                 // 1
    uint64_t* qdo_full = bars + 1;            // [2]
    uint64_t* qdo_empty = bars + 3;           // [2]
    uint64_t* sdp_full = bars + 5;            // S and dP in TMEM (commit)
    uint64_t* sdp_empty = bars + 6;           // 128: S, dP read out of TMEM
    uint64_t* pds_full = bars + 7;            // 128: P, dS written to smem
    uint64_t* pds_empty = bars + 8;           // commit: the three gradient MMAs finished reading P / dS
    uint64_t* dq_full = bars + 9;             // commit
    uint64_t* dq_empty = bars + 10;           // 128
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb = blockIdx.x;                               // key block (early blocks are the heavy ones)
    const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
    const int d = H * 64, ld3 = 3 * d;
    const int k0 = jb * AT_BN;
    const int nq = (T + AT_BM - 1) / AT_BM;
    const int row_base = b * T;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
        mbar_init(sdp_full, 1); mbar_init(sdp_empty, 256);
        mbar_init(pds_full, 256); mbar_init(pds_empty, 1);
        mbar_init(dq_full, 1); mbar_init(dq_empty, 256);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * AT_TILE);
            tma_load_2d(smem + S::oK, &tmQKV, kv_full, d + h * 64, row_base + k0);
            tma_load_2d(smem + S::oV, &tmQKV, kv_full, 2 * d + h * 64, row_base + k0);
            int it = 0;
            for (int i = jb; i < nq; ++i, ++it) {
                const int st = it & 1;
                mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
                uint8_t* sq = smem + S::oQdO + st * 2 * AT_TILE;
                mbar_arrive_expect_tx(&qdo_full[st], 2 * AT_TILE);
                tma_load_2d(sq, &tmQKV, &qdo_full[st], h * 64, row_base + i * AT_BM);
                tma_load_2d(sq + AT_TILE, &tmDO, &qdo_full[st], h * 64, row_base + i * AT_BM);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);    // S, dP
            constexpr uint32_t idesc_g = make_idesc_bf16(128, 64, true, true);       // dV, dK
            constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);      // dQ
            const uint32_t sK = smem_u32(smem + S::oK), sV = smem_u32(smem + S::oV);
            const uint32_t sP = smem_u32(smem + S::oP), sDS = smem_u32(smem + S::oDS);
            mbar_wait(kv_full, 0);
            int it = 0;
            for (int i = jb; i < nq; ++i, ++it) {
                const int st = it & 1;
                const uint32_t ph = it & 1 ? 1u : 0u;        // barriers that complete once per iteration: parity = it & 1
                const uint32_t sQ = smem_u32(smem + S::oQdO + st * 2 * AT_TILE), sDO = sQ + AT_TILE;
                mbar_wait(&qdo_full[st], (it >> 1) & 1);
                mbar_wait(sdp_empty, ph ^ 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tS, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tDP, desc_kmajor(sDO, k), desc_kmajor(sV, k), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(sdp_full);
                mbar_wait(pds_full, ph);
                mbar_wait(dq_empty, ph ^ 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_bf16(tDV, desc_mnmajor(sP, k), desc_mnmajor(sDO, k), idesc_g, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_bf16(tDK, desc_mnmajor(sDS, k), desc_mnmajor(sQ, k), idesc_g, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_bf16(tDQ, desc_kmajor(sDS, k), desc_mnmajor(sK, k), idesc_q, k > 0 ? 1u : 0u);
                umma_commit(dq_full);
                umma_commit(pds_empty);
                umma_commit(&qdo_empty[st]);
            }
        }
        __syncwarp();
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                    // column half: keys 0-63 / 64-127 of the tile, 32 of the 64 head dims
        const int r = quad * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * kLog2eF;
        const uint32_t sP = smem_u32(smem + S::oP), sDS = smem_u32(smem + S::oDS);
        const uint32_t t32 = drop.thresh16 << 16;
        int it = 0;
        for (int i = jb; i < nq; ++i, ++it) {
            const uint32_t ph = it & 1 ? 1u : 0u;
            const int qi = i * AT_BM + r;                      // this thread's query
            const bool q_ok = qi < T;
            const float lse2 = q_ok ? lse[(size_t)bh * T + qi] * kLog2eF : 0.f;
            const float dlt = q_ok ? delta[(size_t)bh * T + qi] : 0.f;
            const bool need_mask = (i == jb) || (k0 + AT_BN > T) || (i * AT_BM + AT_BM > T);
            const AttnDropRow rk = attn_drop_row(drop.seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi);
            const int kc0 = k0 + half * 64;
            mbar_wait(sdp_full, ph);
            tc_fence_after();
            mbar_wait(pds_empty, ph ^ 1);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t sv[32], gv[32];
                __syncwarp();
                tmem_ld_32x32(tS + lane_off + half * 64 + c * 32, sv);
                tmem_ld_32x32(tDP + lane_off + half * 64 + c * 32, gv);
                tmem_ld_wait();
                float p[32], ds[32];
                if (need_mask) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int kj = kc0 + c * 32 + e;
                        p[e] = (q_ok && kj <= qi && kj < T) ? ex2_fast(fmaf(__uint_as_float(sv[e]), sl2, -lse2)) : 0.f;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) p[e] = ex2_fast(fmaf(__uint_as_float(sv[e]), sl2, -lse2));
                }
                // The softmax scale (a power of two) and the dropout scale are applied to the OUTPUTS (dQ, dK resp. dV), not per element:
                //   ds = P (mask * dP / (1-p) - delta)      pd = mask * P
                if (drop.thresh16) {
                    const uint32_t g0 = (uint32_t)(kc0 + c * 32) >> 2;
#pragma unroll
                    for (int e4 = 0; e4 < 8; ++e4) {
                        uint32_t w0, w1;
                        attn_drop_words(rk, g0 + e4, w0, w1);
                        const bool k4[4] = {w0 >= t32, (w0 << 16) >= t32, w1 >= t32, (w1 << 16) >= t32};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float u = fmaf(__uint_as_float(gv[e4 * 4 + e]), drop.scale, -dlt);
                            ds[e4 * 4 + e] = p[e4 * 4 + e] * (k4[e] ? u : -dlt);
                            p[e4 * 4 + e] = k4[e] ? p[e4 * 4 + e] : 0.f;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) ds[e] = p[e] * (__uint_as_float(gv[e]) - dlt);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    st_tile_chunk(sP, r, (half * 2 + c) * 4 + g,
                                  make_uint4(pack_bf16(p[8 * g], p[8 * g + 1]), pack_bf16(p[8 * g + 2], p[8 * g + 3]),
                                             pack_bf16(p[8 * g + 4], p[8 * g + 5]), pack_bf16(p[8 * g + 6], p[8 * g + 7])));
                    st_tile_chunk(sDS, r, (half * 2 + c) * 4 + g,
                                  make_uint4(pack_bf16(ds[8 * g], ds[8 * g + 1]), pack_bf16(ds[8 * g + 2], ds[8 * g + 3]),
                                             pack_bf16(ds[8 * g + 4], ds[8 * g + 5]), pack_bf16(ds[8 * g + 6], ds[8 * g + 7])));
                }
            }
            tc_fence_before();
            mbar_arrive(sdp_empty);
            fence_proxy_async();
            mbar_arrive(pds_full);
            // dQ tile of this (query block, key block) pair -> fp32 accumulation buffer (this thread: 32 of the 64 dims)
            mbar_wait(dq_full, ph);
            tc_fence_after();
            float* dst = dq_acc + (size_t)(row_base + qi) * d + h * 64 + half * 32;
            {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(tDQ + lane_off + half * 32, v);
                tmem_ld_wait();
                if (q_ok) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * g), "f"(__uint_as_float(v[4 * g])),
                                     "f"(__uint_as_float(v[4 * g + 1])), "f"(__uint_as_float(v[4 * g + 2])), "f"(__uint_as_float(v[4 * g + 3])) : "memory");
                }
            }
            tc_fence_before();
            mbar_arrive(dq_empty);
        }
        // dK, dV of this key block (complete once the last iteration's commit has fired: dq_full of that iteration)
        const int kj = k0 + r;
        bf16* dkp = dqkv + (size_t)(row_base + min(kj, T - 1)) * ld3 + d + h * 64 + half * 32;
        bf16* dvp = dkp + d;
        {
            uint32_t a[32], v[32];
            __syncwarp();                                   // .aligned TMEM loads: whole warp, unconditionally
            tmem_ld_32x32(tDK + lane_off + half * 32, a);
            tmem_ld_32x32(tDV + lane_off + half * 32, v);
            tmem_ld_wait();
            if (kj < T) {
#pragma unroll
                for (int e = 0; e < 32; ++e) { a[e] = __float_as_uint(__uint_as_float(a[e]) * scale); v[e] = __float_as_uint(__uint_as_float(v[e]) * drop.scale); }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    reinterpret_cast<uint4*>(dkp)[g] =
                        make_uint4(pack_bf16(__uint_as_float(a[8 * g]), __uint_as_float(a[8 * g + 1])), pack_bf16(__uint_as_float(a[8 * g + 2]), __uint_as_float(a[8 * g + 3])),
                                   pack_bf16(__uint_as_float(a[8 * g + 4]), __uint_as_float(a[8 * g + 5])), pack_bf16(__uint_as_float(a[8 * g + 6]), __uint_as_float(a[8 * g + 7])));
                    reinterpret_cast<uint4*>(dvp)[g] =
                        make_uint4(pack_bf16(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])), pack_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                   pack_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), pack_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// fp32 dQ accumulation buffer [B*T, d] -> bf16 dqkv[:, 0:d]
__global__ void attn_dq_convert_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv, size_t rows, int d, float scale) {
    const size_t n4 = rows * (size_t)d / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;
        const size_t row = e / d;
        const int c = (int)(e - row * d);
        const float4 v = reinterpret_cast<const float4*>(dq_acc)[i];
        *reinterpret_cast<uint2*>(dqkv + row * 3 * d + c) = make_uint2(pack_bf16(v.x * scale, v.y * scale), pack_bf16(v.z * scale, v.w * scale));
    }
}

bool attn_use_tc() {
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("TTTS_ATTN_LEGACY"); legacy = (e && e[0] == '1') ? 1 : 0; }
    return !legacy;
}

int attn_fwd_tc(const bf16* qkv, bf16* o, float* lse, int B, int T, int H, DropCfg drop, cudaStream_t st) {
    const int d = H * 64;
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, qkv, 2, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, 128, true);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kBytes)); attr = true; }
    dim3 grid((T + AT_BM - 1) / AT_BM, B * H);
    attn_fwd_tc_kernel<<<grid, AT_THREADS, FwdSmem::kBytes, st>>>(tm, o, lse, T, H, 0.125f, drop);
    TTTS_LAUNCH_CHECK("attn_fwd_tc");
    return TTTS_OK;
}

// defined in attention.cu
int attn_delta(const bf16* o, const bf16* dout, float* delta, int B, int T, int H, cudaStream_t st);

int attn_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int B, int T, int H, DropCfg drop,
                cudaStream_t st) {
    const int d = H * 64;
    float* dq_acc = delta + (((size_t)B * H * T + 63) / 64) * 64;       // scratch layout: delta | dq_acc[B*T, d]
    TTTS_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * T * d * sizeof(float), st));
    int rc = attn_delta(o, dout, delta, B, T, H, st);
    if (rc) return rc;
    CUtensorMap tmQ, tmDO;
    rc = make_tmap_2d(&tmQ, qkv, 2, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, 128, true);
    if (rc) return rc;
    rc = make_tmap_2d(&tmDO, dout, 2, (uint64_t)d, (uint64_t)B * T, (uint64_t)d, 64, 128, true);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kBytes)); attr = true; }
    dim3 grid((T + AT_BN - 1) / AT_BN, B * H);
    attn_bwd_tc_kernel<<<grid, AT_THREADS, BwdSmem::kBytes, st>>>(tmQ, tmDO, lse, delta, dqkv, dq_acc, T, H, 0.125f, drop);
    TTTS_LAUNCH_CHECK("attn_bwd_tc");
    const size_t n4 = (size_t)B * T * d / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    attn_dq_convert_kernel<<<blocks, 256, 0, st>>>(dq_acc, dqkv, (size_t)B * T, d, 0.125f);     // dQ = scale * dS K
    TTTS_LAUNCH_CHECK("attn_dq_convert");
    return TTTS_OK;
}

}  // namespace ttts
