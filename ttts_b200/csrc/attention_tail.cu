// The ragged tail of causal attention: the last T mod 128 query rows of every (sequence, head), on CUDA cores.
//
// Why: the tcgen05 kernels (attention_tc.cu) work on 128-query x 128-key tiles.  The reference's training sequence is
// text 128 + 2, codes 1024 + 2 = 1156 = 9 x 128 + 4 positions (ttts/gpt/model.py:454-489: start / stop tokens around both segments), so a
// tenth query tile with FOUR valid rows walked all ten key blocks at full tile cost: 10 of 55 block pairs per head = 18 % of the forward and of
// the backward (cfg2: T = 644 = 5 x 128 + 4, 6 of 21 pairs).  When 0 < T mod 128 <= 16 the tile kernels now stop at Tm = T - T mod 128 (all
// their tiles full, no sequence-end masks) and these two kernels do the r = T - Tm tail rows: r x T x 64 multiply-adds per head.
//
// STATUS: opt-in (TTTS_ATTN_TAIL=1), parity-green on a B200 (r2ab, r2ac) but NOT a win: the arithmetic is negligible, the traffic is not --
// four query rows meet every key, so the forward streams K and V (151 MB at B = 32, H = 16, T = 1156) once more and the backward adds a
// read-modify-write of every dK / dV row (453 MB, 70 us at the HBM roof): 61 / 200 us measured against 51 / 90 us saved in the tile kernels.
// Kept as the reference for the in-tile version (tail rows handled while the key / value tiles sit in shared memory), which removes that traffic.
//
// Same arithmetic and rounding points as the tile kernels (HF: modeling_gpt2.py:185-226; dropout on the probabilities, :207-222):
//   forward : s = q . k (bf16 products, fp32 sum) ; m = max s ; p = 2^(s c - m c), c = scale log2 e ; l = sum p (fp32, before dropout) ;
//             out = bf16( sum_j keep_j bf16(p_j) v_j * (1 / (1 - p_drop)) / l ) ; lse = m scale + ln l
//   backward: pe = 2^(s c - lse log2 e) ; dP = dO . v ; nd = pe (-delta (1 - p_drop)) ; dS~ = keep ? bf16(fma(pe, dP, nd)) : bf16(nd) ;
//             P~ = keep ? bf16(pe) : 0 ;  dQ_acc[i] = sum_j dS~_ij k_j (fp32; attn_dq_convert scales it) ;
//             dK_j += scale / (1 - p_drop) sum_i dS~_ij q_i ;  dV_j += 1 / (1 - p_drop) sum_i P~_ij dO_i
//   The keep decisions are the same hash of (seed, (b H + h) T + query, key) as everywhere else (attn_drop_hash.h).
// The tail rows' dK / dV contributions are ADDED to the bf16 rows the tile kernel has already written for keys < Tm (read - add in fp32 -
// round - write; this kernel runs after it on the same stream) and stored directly for the keys >= Tm, which only tail queries see.
//
// One CTA per (sequence, head), 256 threads; shared memory holds the r x T score / probability rows (fp32).  Also compiled for the HOST by
// tests/emu (g++ -DTTTS_HOST_EMU), which runs these kernels on CPU against torch: hence no <<<>>> and the TTTS_DYN_SMEM macro.
#include <math.h>
#include <string.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#include "../../ttts_b200/csrc/attn_drop_hash.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#define TTTS_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

namespace ttts {

constexpr int AT_TAIL_MAX = 16;           // largest T mod 128 handled here
constexpr int AT_TAIL_THREADS = 256;
constexpr float kTailLog2e = 1.4426950408889634f;

TTTS_DEVICE float tail_bf(const bf16* p) { return __bfloat162float(*p); }
// 8 consecutive bf16 (16 bytes, 16-byte aligned) -> 8 floats
TTTS_DEVICE void tail_ld8(const bf16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
TTTS_DEVICE bool tail_keep(const AttnDropRow rk, int j, uint32_t thresh16) {
    if (!thresh16) return true;
    uint32_t w0, w1;
    attn_drop_words(rk, (uint32_t)j >> 2, w0, w1);
    return attn_drop_keep(w0, w1, j & 3, thresh16 << 16);
}

TTTS_DEVICE void tail_unpack8(const uint4 v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

// red[warp][i][c] = this warp's share of sum_j w[i][j] m[j][c]   (w: shared [R][T]; m: bf16 rows of 64 dims, row pitch ld elements).
// lane = (key mod 4, 8-dim chunk): a warp load covers four whole 128-byte rows; warp x takes keys 4 x .. 4 x + 3 of every 32.  The loads of
// four rounds are requested before the first multiply-add: r2ab measured the one-load-per-iteration form of this loop at ~50 us per CTA
// (every iteration exposed a full L2 round trip), slower than the tile it was meant to replace.
template <int R>
TTTS_DEVICE void tail_rows_times_matrix(const float* w, const bf16* mp, int ld, int T, float* red) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kq = lane >> 3, c8 = lane & 7;
    float acc[R][8];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
    for (int j0 = warp * 4 + kq; j0 < T; j0 += 32 * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + 32 * u;
            raw[u] = j < T ? *reinterpret_cast<const uint4*>(mp + (size_t)j * ld + c8 * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + 32 * u;
            if (j < T) {
                float mf[8];
                tail_unpack8(raw[u], mf);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const float wv = w[(size_t)i * T + j];
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(wv, mf[e], acc[i][e]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = acc[i][e];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (kq == 0) red[(warp * R + i) * 64 + c8 * 8 + e] = v;
        }
}
TTTS_DEVICE float tail_red8(const float* red, int R, int i, int c) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < AT_TAIL_THREADS / 32; ++w) v += red[(w * R + i) * 64 + c];
    return v;
}

// shared memory (floats): q rows [R][64] | (backward: dO rows [R][64]) | row stats [2][R] | score rows [R][T] (backward: two of them) |
// cross-warp reduction [8][R][64]
template <int R>
__global__ void __launch_bounds__(AT_TAIL_THREADS, R <= 4 ? 4 : (R <= 8 ? 2 : 1))
attn_tail_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse_out, int T, int Tm, int H, float scale,
                     uint32_t thresh16, float drop_scale, uint64_t seed) {
    TTTS_DYN_SMEM(float, tail_sm);
    float* qs = tail_sm;                      // [R][64]
    float* stat = qs + R * 64;                // [R] row sums
    float* sc = stat + 2 * R;                 // [R][T]
    float* red = sc + (size_t)R * T;          // [8][R][64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H, d = H * 64, ld = 3 * d;
    const int r = T - Tm;
    const float sl2 = scale * kTailLog2e;
    const bf16* base = qkv + (size_t)b * T * ld + h * 64;
    pdl_launch_dependents();
    pdl_wait();
    for (int e = tid; e < R * 64; e += AT_TAIL_THREADS) {
        const int i = e >> 6, c = e & 63;
        qs[e] = i < r ? tail_bf(base + (size_t)(Tm + i) * ld + c) : 0.f;
    }
    __syncthreads();
    // scores: one key per thread, all tail rows at once
    for (int j = tid; j < T; j += AT_TAIL_THREADS) {
        const bf16* kp = base + (size_t)j * ld + d;
        float acc[R];
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float kf[8];
            tail_ld8(kp + c8 * 8, kf);
#pragma unroll
            for (int i = 0; i < R; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[i] = fmaf(kf[e], qs[i * 64 + c8 * 8 + e], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < R; ++i) sc[(size_t)i * T + j] = acc[i];
    }
    __syncthreads();
    // softmax of row i over its Tm + i + 1 keys: one warp per row
    for (int i = warp; i < R; i += AT_TAIL_THREADS / 32) {
        float* row = sc + (size_t)i * T;
        if (i >= r) {
            for (int j = lane; j < T; j += 32) row[j] = 0.f;
            continue;
        }
        const int qi = Tm + i, n = qi + 1;
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        const float msc = m * sl2;
        const AttnDropRow rk = attn_drop_row(seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi);
        float l = 0.f;
        for (int j = lane; j < T; j += 32) {
            float pk = 0.f;
            if (j < n) {
                const float p = exp2f(fmaf(row[j], sl2, -msc));
                l += p;
                pk = tail_keep(rk, j, thresh16) ? bf16_round(p) : 0.f;
            }
            row[j] = pk;
        }
        l = warp_sum(l);
        if (lane == 0) {
            stat[i] = l;
            lse_out[(size_t)bh * T + qi] = m * scale + logf(l);
        }
    }
    __syncthreads();
    // out rows = P V
    tail_rows_times_matrix<R>(sc, base + 2 * d, ld, T, red);
    __syncthreads();
    for (int e = tid; e < r * 64; e += AT_TAIL_THREADS) {
        const int i = e >> 6, c = e & 63;
        const float o = tail_red8(red, R, i, c);
        const float l = stat[i];
        const float inv = l > 0.f ? drop_scale / l : 0.f;
        out[(size_t)(b * T + Tm + i) * d + h * 64 + c] = __float2bfloat16_rn(o * inv);
    }
}

template <int R>
__global__ void __launch_bounds__(AT_TAIL_THREADS, R <= 4 ? 4 : (R <= 8 ? 2 : 1))
attn_tail_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                     bf16* __restrict__ dqkv, float* __restrict__ dq_acc, int T, int Tm, int H, float scale, uint32_t thresh16, float drop_scale,
                     uint64_t seed) {
    TTTS_DYN_SMEM(float, tail_sm);
    float* qs = tail_sm;                      // [R][64]
    float* dos = qs + R * 64;                 // [R][64]
    float* stat = dos + R * 64;               // [2][R]: lse log2 e | -delta / c  (the forward layout keeps 2 R floats here too), then
    uint32_t* rkw = reinterpret_cast<uint32_t*>(stat + 2 * R);      // [2][R]: the rows' dropout keys
    float* sp = stat + 4 * R;                 // [R][T]  P~
    float* sd = sp + (size_t)R * T;           // [R][T]  dS~
    float* red = sd + (size_t)R * T;          // [8][R][64]
    const int tid = threadIdx.x;
    const int bh = blockIdx.x, b = bh / H, h = bh - b * H, d = H * 64, ld = 3 * d;
    const int r = T - Tm;
    const float sl2 = scale * kTailLog2e;
    const bf16* base = qkv + (size_t)b * T * ld + h * 64;
    pdl_launch_dependents();
    pdl_wait();
    for (int e = tid; e < R * 64; e += AT_TAIL_THREADS) {
        const int i = e >> 6, c = e & 63;
        qs[e] = i < r ? tail_bf(base + (size_t)(Tm + i) * ld + c) : 0.f;
        dos[e] = i < r ? tail_bf(dout + (size_t)(b * T + Tm + i) * d + h * 64 + c) : 0.f;
    }
    if (tid < R) {
        const bool ok = tid < r;
        stat[tid] = ok ? lse[(size_t)bh * T + Tm + tid] * kTailLog2e : 0.f;
        stat[R + tid] = ok ? -delta[(size_t)bh * T + Tm + tid] / drop_scale : 0.f;
        const AttnDropRow rk = attn_drop_row(seed, (uint64_t)bh * (uint64_t)T + (uint64_t)(Tm + tid));
        rkw[tid] = rk.k0; rkw[R + tid] = rk.k1;
    }
    __syncthreads();
    // P~ and dS~ of every (tail row, key): one key per thread
    for (int j = tid; j < T; j += AT_TAIL_THREADS) {
        const bf16* kp = base + (size_t)j * ld + d;
        const bf16* vp = kp + d;
        float s[R], dp[R];
#pragma unroll
        for (int i = 0; i < R; ++i) { s[i] = 0.f; dp[i] = 0.f; }
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float kf[8], vf[8];
            tail_ld8(kp + c8 * 8, kf);
            tail_ld8(vp + c8 * 8, vf);
#pragma unroll
            for (int i = 0; i < R; ++i)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s[i] = fmaf(kf[e], qs[i * 64 + c8 * 8 + e], s[i]);
                    dp[i] = fmaf(vf[e], dos[i * 64 + c8 * 8 + e], dp[i]);
                }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            float pv = 0.f, dv = 0.f;
            const int qi = Tm + i;
            if (i < r && j <= qi) {
                const float pe = exp2f(fmaf(s[i], sl2, -stat[i]));
                const float nd = pe * stat[R + i];
                const float dk = fmaf(pe, dp[i], nd);
                AttnDropRow rk; rk.k0 = rkw[i]; rk.k1 = rkw[R + i];
                const bool keep = tail_keep(rk, j, thresh16);
                pv = keep ? bf16_round(pe) : 0.f;
                dv = bf16_round(keep ? dk : nd);
            }
            sp[(size_t)i * T + j] = pv;
            sd[(size_t)i * T + j] = dv;
        }
    }
    __syncthreads();
    // dK / dV rows: thread = (key, 8-dim chunk); read - add - write for the keys the tile kernel has written (j < Tm).  Four items per round,
    // all eight loads requested before the first store (a load may not be moved above a store the compiler cannot tell apart from it)
    {
        constexpr int U = 4;
        const float ks = scale * drop_scale;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int it0 = tid; it0 < T * 8; it0 += AT_TAIL_THREADS * U) {
            uint4 rk4[U], rv4[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int it = it0 + AT_TAIL_THREADS * u, j = it >> 3, c8 = it & 7;
                const bool rd = it < T * 8 && j < Tm;
                const bf16* dkp = dqkv + (size_t)(b * T + (rd ? j : 0)) * ld + d + h * 64 + c8 * 8;
                rk4[u] = rd ? *reinterpret_cast<const uint4*>(dkp) : z4;
                rv4[u] = rd ? *reinterpret_cast<const uint4*>(dkp + d) : z4;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int it = it0 + AT_TAIL_THREADS * u, j = it >> 3, c8 = it & 7;
                if (it < T * 8) {
                    float ak[8], av[8], ok[8], ov[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { ak[e] = 0.f; av[e] = 0.f; }
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const float ds = sd[(size_t)i * T + j], p = sp[(size_t)i * T + j];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            ak[e] = fmaf(ds, qs[i * 64 + c8 * 8 + e], ak[e]);
                            av[e] = fmaf(p, dos[i * 64 + c8 * 8 + e], av[e]);
                        }
                    }
                    tail_unpack8(rk4[u], ok);
                    tail_unpack8(rv4[u], ov);
#pragma unroll
                    for (int e = 0; e < 8; ++e) { ok[e] = fmaf(ak[e], ks, ok[e]); ov[e] = fmaf(av[e], drop_scale, ov[e]); }
                    bf16* dkp = dqkv + (size_t)(b * T + j) * ld + d + h * 64 + c8 * 8;
                    *reinterpret_cast<uint4*>(dkp) = make_uint4(pack_bf16(ok[0], ok[1]), pack_bf16(ok[2], ok[3]), pack_bf16(ok[4], ok[5]), pack_bf16(ok[6], ok[7]));
                    *reinterpret_cast<uint4*>(dkp + d) = make_uint4(pack_bf16(ov[0], ov[1]), pack_bf16(ov[2], ov[3]), pack_bf16(ov[4], ov[5]), pack_bf16(ov[6], ov[7]));
                }
            }
        }
    }
    // dQ rows (fp32 accumulator, unscaled) = dS~ K
    tail_rows_times_matrix<R>(sd, base + d, ld, T, red);
    __syncthreads();
    for (int e = tid; e < r * 64; e += AT_TAIL_THREADS) {
        const int i = e >> 6, c = e & 63;
        dq_acc[(size_t)(b * T + Tm + i) * d + h * 64 + c] = tail_red8(red, R, i, c);
    }
}

static size_t tail_smem_bytes(int R, int T, bool bwd) {
    return sizeof(float) * ((size_t)(bwd ? 2 : 1) * R * 64 + (bwd ? 4 : 2) * R + (size_t)(bwd ? 2 : 1) * R * T + (AT_TAIL_THREADS / 32) * R * 64);
}
static int tail_R(int r) { return r <= 4 ? 4 : (r <= 8 ? 8 : 16); }

// number of tail rows the tile kernels should leave to this file (0: no split)
int attn_tail_rows(int T) {
    const int r = T % 128;
    if (T <= 128 || r == 0 || r > AT_TAIL_MAX) return 0;
    if (tail_smem_bytes(tail_R(r), T, true) > 200 * 1024) return 0;
    return r;
}

int attn_tail_fwd(const bf16* qkv, bf16* out, float* lse, int B, int T, int H, int Tm, uint32_t thresh16, float drop_scale, uint64_t seed,
                  cudaStream_t st) {
    const int r = T - Tm;
    TTTS_CHECK_ARG(qkv && out && lse && B >= 1 && H >= 1 && Tm >= 1 && r >= 1 && r <= AT_TAIL_MAX, "attn_tail_fwd: bad arguments");
    const int R = tail_R(r);
    const size_t smem = tail_smem_bytes(R, T, false);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "attn_tail_fwd: sequence too long for the shared-memory score rows");
#define TTTS_TAIL_FWD(RR)                                                                                                                     \
    do {                                                                                                                                      \
        TTTS_CUDA(cudaFuncSetAttribute(attn_tail_fwd_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        TTTS_CUDA(launch_pdl(attn_tail_fwd_kernel<RR>, dim3(B * H), dim3(AT_TAIL_THREADS), smem, st, qkv, out, lse, T, Tm, H, 0.125f,         \
                             thresh16, drop_scale, seed));                                                                                    \
    } while (0)
    if (R == 4) TTTS_TAIL_FWD(4); else if (R == 8) TTTS_TAIL_FWD(8); else TTTS_TAIL_FWD(16);
#undef TTTS_TAIL_FWD
    TTTS_LAUNCH_CHECK("attn_tail_fwd");
    return TTTS_OK;
}

int attn_tail_bwd(const bf16* qkv, const bf16* dout, const float* lse, const float* delta, bf16* dqkv, float* dq_acc, int B, int T, int H, int Tm,
                  uint32_t thresh16, float drop_scale, uint64_t seed, cudaStream_t st) {
    const int r = T - Tm;
    TTTS_CHECK_ARG(qkv && dout && lse && delta && dqkv && dq_acc && B >= 1 && H >= 1 && Tm >= 1 && r >= 1 && r <= AT_TAIL_MAX,
                   "attn_tail_bwd: bad arguments");
    const int R = tail_R(r);
    const size_t smem = tail_smem_bytes(R, T, true);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "attn_tail_bwd: sequence too long for the shared-memory score rows");
#define TTTS_TAIL_BWD(RR)                                                                                                                     \
    do {                                                                                                                                      \
        TTTS_CUDA(cudaFuncSetAttribute(attn_tail_bwd_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        TTTS_CUDA(launch_pdl(attn_tail_bwd_kernel<RR>, dim3(B * H), dim3(AT_TAIL_THREADS), smem, st, qkv, dout, lse, delta, dqkv, dq_acc, T,  \
                             Tm, H, 0.125f, thresh16, drop_scale, seed));                                                                     \
    } while (0)
    if (R == 4) TTTS_TAIL_BWD(4); else if (R == 8) TTTS_TAIL_BWD(8); else TTTS_TAIL_BWD(16);
#undef TTTS_TAIL_BWD
    TTTS_LAUNCH_CHECK("attn_tail_bwd");
    return TTTS_OK;
}

}  // namespace ttts
