// Kernels the diffusion mel-refiner train step adds to the training tape (BASELINE config 5;
// ttts_b200/diffusion/train_graph.py).  All tensors fp32, [B, C, T] channel-major like the convolution kernels.
//   * GroupNorm32 (+ the ResBlock's (1 + scale) / shift modulation, + SiLU) forward / backward        ttts/utils/utils.py:119-137, aa_model.py:120-135
//   * SiLU
//   * QKVAttentionLegacy + RelativePositionBias: non-causal attention over the packed head-major qkv with a bucketed relative-position
//     bias, flash-style (no [T, T] tensor in memory), forward and backward                            utils.py:148-175, xtransformers.py:146-188
//   * q_sample and the training loss (MSE + learned-range variational-bound term)                      utils/diffusion.py:243-260, 903-1014
// Plain CUDA (no TMA / tcgen05), exact fp32 like the reference's default (no autocast) trainer; the CPU emulation of this source
// (tests/emu/diffusion_emu.cpp) is checked against the op contract tests/ref_kernels.py, the GPU tests against the same contract.
#include <stdlib.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#define TTTS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace ttts {

TTTS_DEVICE float dsig(float x) { return 1.f / (1.f + expf(-x)); }

// sum over a 256-thread block (8 warps), result broadcast to every thread; `red` = 8 floats of shared memory
TTTS_DEVICE float block_sum256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    return s;
}

// ---------------------------------------------------------------- GroupNorm ----------------------------------------------------------------
// one CTA per (group, batch): the group's cpg * T floats are contiguous.  Two-pass statistics (mean, then centred second moment), then
// y = act((xhat * gamma + beta) * (1 + scale[b, c]) + shift[b, c]).  stats [B, G, 2] = (mean, rstd) kept for the backward.
__global__ void __launch_bounds__(256) gn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                                                     float* __restrict__ stats, int C, int T, int G, int silu) {
    __shared__ float red[8];
    const int g = blockIdx.x, b = blockIdx.y, cpg = C / G;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const int n = cpg * T;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += x[base + i];
    const float mean = block_sum256(s, red) / (float)n;
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) { const float d = x[base + i] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(block_sum256(v, red) / (float)n + 1e-5f);
    if (threadIdx.x == 0) { stats[((size_t)b * G + g) * 2] = mean; stats[((size_t)b * G + g) * 2 + 1] = rstd; }
    for (int i = threadIdx.x; i < n; i += 256) {
        const int c = g * cpg + i / T;
        float u = (x[base + i] - mean) * rstd * gamma[c] + beta[c];
        if (scale) u = u * (1.f + scale[(size_t)b * C + c]) + shift[(size_t)b * C + c];
        y[base + i] = silu ? u * dsig(u) : u;
    }
}

// backward: one CTA per (group, batch).  Pass 1: per channel of the group the sums over time of dm = dy act'(m), dm u, du = dm (1 + scale)
// and du xhat (one warp per channel, fixed order); pass 2: dx = rstd (du gamma - mean_grp(du gamma) - xhat mean_grp(du gamma xhat)).
// part [B, C, 2] = per-(b, c) (sum du xhat, sum du): reduced over the batch by gn_param_kernel (deterministic).
__global__ void __launch_bounds__(256) gn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, float* __restrict__ dx, float* __restrict__ part,
                                                     float* __restrict__ dscale, float* __restrict__ dshift, int C, int T, int G, int silu) {
    __shared__ float sums[64][2];
    const int g = blockIdx.x, b = blockIdx.y, cpg = C / G;
    const size_t base = ((size_t)b * C + (size_t)g * cpg) * T;
    const float mean = stats[((size_t)b * G + g) * 2], rstd = stats[((size_t)b * G + g) * 2 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cl = warp; cl < cpg; cl += 8) {
        const int c = g * cpg + cl;
        const float ga = gamma[c], be = beta[c];
        const float sc = scale ? 1.f + scale[(size_t)b * C + c] : 1.f, sh = scale ? shift[(size_t)b * C + c] : 0.f;
        float s_dm = 0.f, s_dmu = 0.f, s_dux = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float xh = (x[base + (size_t)cl * T + t] - mean) * rstd;
            const float u = xh * ga + be, m = u * sc + sh;
            float d = dy[base + (size_t)cl * T + t];
            if (silu) { const float sg = dsig(m); d *= sg * (1.f + m * (1.f - sg)); }
            s_dm += d; s_dmu = fmaf(d, u, s_dmu); s_dux = fmaf(d * sc, xh, s_dux);
        }
        s_dm = warp_sum(s_dm); s_dmu = warp_sum(s_dmu); s_dux = warp_sum(s_dux);
        if (lane == 0) {
            sums[cl][0] = s_dm * sc; sums[cl][1] = s_dux;
            part[((size_t)b * C + c) * 2] = s_dux; part[((size_t)b * C + c) * 2 + 1] = s_dm * sc;
            if (dscale) { dscale[(size_t)b * C + c] = s_dmu; dshift[(size_t)b * C + c] = s_dm; }
        }
    }
    __syncthreads();
    float A = 0.f, Bq = 0.f;
    for (int cl = 0; cl < cpg; ++cl) { const float ga = gamma[g * cpg + cl]; A = fmaf(ga, sums[cl][0], A); Bq = fmaf(ga, sums[cl][1], Bq); }
    const int n = cpg * T;
    A /= (float)n; Bq /= (float)n;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int c = g * cpg + i / T;
        const float ga = gamma[c];
        const float sc = scale ? 1.f + scale[(size_t)b * C + c] : 1.f, sh = scale ? shift[(size_t)b * C + c] : 0.f;
        const float xh = (x[base + i] - mean) * rstd;
        float d = dy[base + i];
        if (silu) { const float m = (xh * ga + beta[c]) * sc + sh; const float sg = dsig(m); d *= sg * (1.f + m * (1.f - sg)); }
        dx[base + i] = rstd * (d * sc * ga - A - xh * Bq);
    }
}
// dgamma[c] = sum_b part[b, c, 0], dbeta[c] = sum_b part[b, c, 1]
__global__ void gn_param_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, e = 0.f;
    for (int b = 0; b < B; ++b) { a += part[((size_t)b * C + c) * 2]; e += part[((size_t)b * C + c) * 2 + 1]; }
    dgamma[c] = a; dbeta[c] = e;
}

__global__ void silu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ o, size_t n, int dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i], sg = dsig(v);
        o[i] = dir == 0 ? v * sg : dy[i] * sg * (1.f + v * (1.f - sg));
    }
}

// ---------------------------------------------------------------- attention with relative-position bias ----------------------------------------------------------------
// qkv [B, 3C, T]: head h owns channels [3 ch h, 3 ch (h+1)) = q | k | v blocks of ch channels (QKVAttentionLegacy splits the heads first);
// scores[i, j] = q_i . k_j / sqrt(ch) + sqrt(ch) table[diag[j - i + T - 1], h]; softmax over j; out [B, C, T] channel h ch + d.
// 64 x 64 tiles, 256 threads as a 16 x 16 grid (ty = tid / 16 owns rows 4 ty .. 4 ty + 3, tx = tid % 16 owns columns 4 tx .. 4 tx + 3 of a
// score tile and dims tx, tx + 16, ... of an output tile); operand tiles live in shared memory d-major [ch][PITCH] exactly as they lie in
// global memory (coalesced along time), score tiles as [64][PITCH].
constexpr int AT = 64, PITCH = 68;

struct AttnBiasParams {
    const float *qkv, *table, *out, *dout, *lse_in, *delta;
    const int* diag;
    float *o, *lse, *dqkv, *dpart;
    int C, T, H, ch;
};

// tile [ch][PITCH] <- rows [row0, row0 + ch) of a [., T] matrix, columns [t0, t0 + 64), zero beyond T, times mul
TTTS_DEVICE void load_tile(float* s, const float* g, int ch, int T, int t0, float mul) {
    for (int i = threadIdx.x; i < ch * AT; i += 256) {
        const int d = i >> 6, t = i & 63;
        s[d * PITCH + t] = (t0 + t < T) ? g[(size_t)d * T + t0 + t] * mul : 0.f;
    }
}
// s[ii][jj] = sum_d A[d][4 ty + ii] B[d][4 tx + jj]
TTTS_DEVICE void tile_nn(const float* A, const float* Bm, int ch, int ty, int tx, float s[4][4]) {
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) s[ii][jj] = 0.f;
    for (int d = 0; d < ch; ++d) {
        const float4 a = *reinterpret_cast<const float4*>(A + d * PITCH + 4 * ty);
        const float4 b = *reinterpret_cast<const float4*>(Bm + d * PITCH + 4 * tx);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) s[ii][jj] = fmaf(av[ii], bv[jj], s[ii][jj]);
    }
}
// acc[ii][dd] += sum_j M[4 ty + ii][j] X[tx + 16 dd][j]        (M [64][PITCH], X [ch][PITCH]; reduction along the rows of both)
template <int DPT>
TTTS_DEVICE void acc_rows(const float* M, const float* X, int ch, int ty, int tx, float acc[4][DPT]) {
    for (int j = 0; j < AT; j += 4) {
        float4 m[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) m[ii] = *reinterpret_cast<const float4*>(M + (4 * ty + ii) * PITCH + j);
#pragma unroll
        for (int dd = 0; dd < DPT; ++dd) {
            const int d = tx + 16 * dd;
            if (d >= ch) break;
            const float4 xv = *reinterpret_cast<const float4*>(X + d * PITCH + j);
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
                acc[ii][dd] = fmaf(m[ii].x, xv.x, fmaf(m[ii].y, xv.y, fmaf(m[ii].z, xv.z, fmaf(m[ii].w, xv.w, acc[ii][dd]))));
        }
    }
}
// acc[jj][dd] += sum_i M[i][4 ty + jj] X[tx + 16 dd][i]        (reduction down the columns of M)
template <int DPT>
TTTS_DEVICE void acc_cols(const float* M, const float* X, int ch, int ty, int tx, float acc[4][DPT]) {
    for (int i = 0; i < AT; i += 4) {
        float4 m[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) m[r] = *reinterpret_cast<const float4*>(M + (i + r) * PITCH + 4 * ty);
#pragma unroll
        for (int dd = 0; dd < DPT; ++dd) {
            const int d = tx + 16 * dd;
            if (d >= ch) break;
            const float4 xv = *reinterpret_cast<const float4*>(X + d * PITCH + i);
            acc[0][dd] = fmaf(m[0].x, xv.x, fmaf(m[1].x, xv.y, fmaf(m[2].x, xv.z, fmaf(m[3].x, xv.w, acc[0][dd]))));
            acc[1][dd] = fmaf(m[0].y, xv.x, fmaf(m[1].y, xv.y, fmaf(m[2].y, xv.z, fmaf(m[3].y, xv.w, acc[1][dd]))));
            acc[2][dd] = fmaf(m[0].z, xv.x, fmaf(m[1].z, xv.y, fmaf(m[2].z, xv.z, fmaf(m[3].z, xv.w, acc[2][dd]))));
            acc[3][dd] = fmaf(m[0].w, xv.x, fmaf(m[1].w, xv.y, fmaf(m[2].w, xv.z, fmaf(m[3].w, xv.w, acc[3][dd]))));
        }
    }
}
// reductions across the 16 threads (tx) that share a row: they are 16 consecutive lanes of one warp
TTTS_DEVICE float row_max16(float v) { for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
TTTS_DEVICE float row_sum16(float v) { for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
// bias of every diagonal of this head: sB[r] = sqrt(ch) table[diag[r], h], r = (j - i) + T - 1
TTTS_DEVICE void load_bias(float* sB, const AttnBiasParams& p, int h) {
    const float sc = sqrtf((float)p.ch);
    for (int r = threadIdx.x; r < 2 * p.T - 1; r += 256) sB[r] = sc * p.table[p.diag[r] * p.H + h];
}
// write an accumulator tile acc[4][DPT] (rows 4 ty + ii, dims tx + 16 dd) times mul to g rows [., T], columns [t0, t0 + 64), coalesced through st
template <int DPT>
TTTS_DEVICE void store_tile(float* st, float* g, const float acc[4][DPT], const float mul[4], int ch, int T, int t0, int ty, int tx) {
    __syncthreads();
#pragma unroll
    for (int dd = 0; dd < DPT; ++dd) {
        const int d = tx + 16 * dd;
        if (d >= ch) break;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) st[d * PITCH + 4 * ty + ii] = acc[ii][dd] * mul[ii];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ch * AT; i += 256) {
        const int d = i >> 6, t = i & 63;
        if (t0 + t < T) g[(size_t)d * T + t0 + t] = st[d * PITCH + t];
    }
}

template <int DPT>
__global__ void __launch_bounds__(256) attn_bias_fwd_kernel(const AttnBiasParams p) {
    TTTS_DYN_SMEM(float, sm);
    const int ch = p.ch, T = p.T, q0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
    float *Qs = sm, *Ks = Qs + ch * PITCH, *Vs = Ks + ch * PITCH, *Ps = Vs + ch * PITCH, *sB = Ps + AT * PITCH;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const float* qg = p.qkv + ((size_t)b * 3 * p.C + (size_t)h * 3 * ch) * T;
    load_tile(Qs, qg, ch, T, q0, rsqrtf((float)ch));
    load_bias(sB, p, h);
    float m[4], l[4], o[4][DPT];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        m[ii] = -INFINITY; l[ii] = 0.f;
#pragma unroll
        for (int dd = 0; dd < DPT; ++dd) o[ii][dd] = 0.f;
    }
    for (int k0 = 0; k0 < T; k0 += AT) {
        __syncthreads();
        load_tile(Ks, qg + (size_t)ch * T, ch, T, k0, 1.f);
        load_tile(Vs, qg + (size_t)2 * ch * T, ch, T, k0, 1.f);
        __syncthreads();
        float s[4][4];
        tile_nn(Qs, Ks, ch, ty, tx, s);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = min(q0 + 4 * ty + ii, T - 1);
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = k0 + 4 * tx + jj;
                s[ii][jj] = j < T ? s[ii][jj] + sB[j - i + T - 1] : -INFINITY;
                mx = fmaxf(mx, s[ii][jj]);
            }
            mx = row_max16(mx);
            const float mn = fmaxf(m[ii], mx), alpha = expf(m[ii] - mn);
            float sum = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) { s[ii][jj] = expf(s[ii][jj] - mn); sum += s[ii][jj]; }
            l[ii] = l[ii] * alpha + row_sum16(sum);
            m[ii] = mn;
#pragma unroll
            for (int dd = 0; dd < DPT; ++dd) o[ii][dd] *= alpha;
            *reinterpret_cast<float4*>(Ps + (4 * ty + ii) * PITCH + 4 * tx) = make_float4(s[ii][0], s[ii][1], s[ii][2], s[ii][3]);
        }
        __syncthreads();
        acc_rows<DPT>(Ps, Vs, ch, ty, tx, o);
    }
    float inv[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        inv[ii] = 1.f / l[ii];
        const int i = q0 + 4 * ty + ii;
        if (tx == 0 && i < T) p.lse[((size_t)b * p.H + h) * T + i] = m[ii] + logf(l[ii]);
    }
    store_tile<DPT>(Ks, p.o + ((size_t)b * p.C + (size_t)h * ch) * T, o, inv, ch, T, q0, ty, tx);
}

// delta[b, h, i] = sum_d dout[b, h ch + d, i] out[b, h ch + d, i]
__global__ void attn_bias_delta_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ delta, int C, int T, int H,
                                       size_t n) {
    const int ch = C / H;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % T);
        const size_t bh = idx / T;
        const size_t b = bh / H, h = bh % H;
        const size_t base = (b * C + h * ch) * T + t;
        float s = 0.f;
        for (int d = 0; d < ch; ++d) s = fmaf(dout[base + (size_t)d * T], out[base + (size_t)d * T], s);
        delta[idx] = s;
    }
}

// probabilities and score gradients of one 64 x 64 tile, shared by the two backward kernels:
// p = exp(s + bias - lse_i), ds = p (dp - delta_i), both zero outside [0, T) x [0, T)
TTTS_DEVICE void bwd_tile(const float* Qs, const float* Ks, const float* dOs, const float* Vs, const float* sB, const float* sL, const float* sD,
                          int ch, int T, int q0, int k0, int ty, int tx, float pt[4][4], float ds[4][4]) {
    float dp[4][4];
    tile_nn(Qs, Ks, ch, ty, tx, pt);
    tile_nn(dOs, Vs, ch, ty, tx, dp);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const int i = q0 + 4 * ty + ii;
        const float lse = sL[4 * ty + ii], dl = sD[4 * ty + ii];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = k0 + 4 * tx + jj;
            const bool ok = i < T && j < T;
            const float pr = ok ? expf(pt[ii][jj] + sB[j - i + T - 1] - lse) : 0.f;
            pt[ii][jj] = pr;
            ds[ii][jj] = pr * (dp[ii][jj] - dl);
        }
    }
}

// dQ and the bias-table gradient: one CTA per (query tile, head, batch) walks the key tiles.
// dpart [B * n_qtiles, H, 32]: per-CTA bucket sums of dS (times sqrt(ch)), reduced in fixed order by attn_bias_dtable_kernel.
template <int DPT>
__global__ void __launch_bounds__(256) attn_bias_bwd_dq_kernel(const AttnBiasParams p) {
    TTTS_DYN_SMEM(float, sm);
    const int ch = p.ch, T = p.T, q0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
    float *Qs = sm, *Ks = Qs + ch * PITCH, *Vs = Ks + ch * PITCH, *dOs = Vs + ch * PITCH, *Ss = dOs + ch * PITCH, *sB = Ss + AT * PITCH,
          *sDiag = sB + 2 * T, *sL = sDiag + 2 * T, *sD = sL + AT;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const float* qg = p.qkv + ((size_t)b * 3 * p.C + (size_t)h * 3 * ch) * T;
    const float scale = rsqrtf((float)ch);
    load_tile(Qs, qg, ch, T, q0, scale);
    load_tile(dOs, p.dout + ((size_t)b * p.C + (size_t)h * ch) * T, ch, T, q0, 1.f);
    load_bias(sB, p, h);
    for (int r = threadIdx.x; r < 2 * T - 1; r += 256) sDiag[r] = 0.f;
    if (threadIdx.x < AT) {
        const int i = q0 + threadIdx.x;
        sL[threadIdx.x] = i < T ? p.lse_in[((size_t)b * p.H + h) * T + i] : 0.f;
        sD[threadIdx.x] = i < T ? p.delta[((size_t)b * p.H + h) * T + i] : 0.f;
    }
    float dq[4][DPT];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int dd = 0; dd < DPT; ++dd) dq[ii][dd] = 0.f;
    for (int k0 = 0; k0 < T; k0 += AT) {
        __syncthreads();
        load_tile(Ks, qg + (size_t)ch * T, ch, T, k0, 1.f);
        load_tile(Vs, qg + (size_t)2 * ch * T, ch, T, k0, 1.f);
        __syncthreads();
        float pt[4][4], ds[4][4];
        bwd_tile(Qs, Ks, dOs, Vs, sB, sL, sD, ch, T, q0, k0, ty, tx, pt, ds);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
            *reinterpret_cast<float4*>(Ss + (4 * ty + ii) * PITCH + 4 * tx) = make_float4(ds[ii][0], ds[ii][1], ds[ii][2], ds[ii][3]);
        __syncthreads();
        acc_rows<DPT>(Ss, Ks, ch, ty, tx, dq);
        // diagonal sums of this tile: thread r owns the tile diagonal jt - it = r - 63; within one tile every global diagonal has one owner
        if (threadIdx.x < 2 * AT - 1) {
            const int off = (int)threadIdx.x - (AT - 1);
            float s = 0.f;
            for (int it = max(0, -off); it < min(AT, AT - off); ++it) s += Ss[it * PITCH + it + off];
            const int r = (k0 - q0) + off + T - 1;
            if (r >= 0 && r < 2 * T - 1) sDiag[r] += s;
        }
    }
    const float mul[4] = {scale, scale, scale, scale};
    store_tile<DPT>(Ss, p.dqkv + ((size_t)b * 3 * p.C + (size_t)h * 3 * ch) * T, dq, mul, ch, T, q0, ty, tx);
    if (threadIdx.x < 32) {
        float s = 0.f;
        for (int r = 0; r < 2 * T - 1; ++r)
            if (p.diag[r] == (int)threadIdx.x) s += sDiag[r];
        p.dpart[(((size_t)b * gridDim.x + blockIdx.x) * p.H + h) * 32 + threadIdx.x] = s * sqrtf((float)ch);
    }
}

// dK and dV: one CTA per (key tile, head, batch) walks the query tiles
template <int DPT>
__global__ void __launch_bounds__(256) attn_bias_bwd_dkv_kernel(const AttnBiasParams p) {
    TTTS_DYN_SMEM(float, sm);
    const int ch = p.ch, T = p.T, k0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
    float *Qs = sm, *Ks = Qs + ch * PITCH, *Vs = Ks + ch * PITCH, *dOs = Vs + ch * PITCH, *Ss = dOs + ch * PITCH, *Ps = Ss + AT * PITCH,
          *sB = Ps + AT * PITCH, *sL = sB + 2 * T, *sD = sL + AT;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const float* qg = p.qkv + ((size_t)b * 3 * p.C + (size_t)h * 3 * ch) * T;
    const float scale = rsqrtf((float)ch);
    load_tile(Ks, qg + (size_t)ch * T, ch, T, k0, 1.f);
    load_tile(Vs, qg + (size_t)2 * ch * T, ch, T, k0, 1.f);
    load_bias(sB, p, h);
    float dk[4][DPT], dv[4][DPT];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int dd = 0; dd < DPT; ++dd) { dk[ii][dd] = 0.f; dv[ii][dd] = 0.f; }
    for (int q0 = 0; q0 < T; q0 += AT) {
        __syncthreads();
        load_tile(Qs, qg, ch, T, q0, scale);
        load_tile(dOs, p.dout + ((size_t)b * p.C + (size_t)h * ch) * T, ch, T, q0, 1.f);
        if (threadIdx.x < AT) {
            const int i = q0 + threadIdx.x;
            sL[threadIdx.x] = i < T ? p.lse_in[((size_t)b * p.H + h) * T + i] : 0.f;
            sD[threadIdx.x] = i < T ? p.delta[((size_t)b * p.H + h) * T + i] : 0.f;
        }
        __syncthreads();
        float pt[4][4], ds[4][4];
        // rows of the tile = queries (ty), columns = keys (tx)
        bwd_tile(Qs, Ks, dOs, Vs, sB, sL, sD, ch, T, q0, k0, ty, tx, pt, ds);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            *reinterpret_cast<float4*>(Ps + (4 * ty + ii) * PITCH + 4 * tx) = make_float4(pt[ii][0], pt[ii][1], pt[ii][2], pt[ii][3]);
            *reinterpret_cast<float4*>(Ss + (4 * ty + ii) * PITCH + 4 * tx) = make_float4(ds[ii][0], ds[ii][1], ds[ii][2], ds[ii][3]);
        }
        __syncthreads();
        acc_cols<DPT>(Ps, dOs, ch, ty, tx, dv);
        acc_cols<DPT>(Ss, Qs, ch, ty, tx, dk);
    }
    const float one[4] = {1.f, 1.f, 1.f, 1.f};
    float* dg = p.dqkv + ((size_t)b * 3 * p.C + (size_t)h * 3 * ch) * T;
    store_tile<DPT>(Ss, dg + (size_t)ch * T, dk, one, ch, T, k0, ty, tx);
    store_tile<DPT>(Ps, dg + (size_t)2 * ch * T, dv, one, ch, T, k0, ty, tx);
}
// dtable[bucket, h] = sum over (batch, query tile) of dpart, fixed order
__global__ void attn_bias_dtable_kernel(const float* __restrict__ dpart, float* __restrict__ dtable, int n_parts, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;           // = h * 32 + bucket
    if (i >= H * 32) return;
    const int h = i / 32, k = i % 32;
    float s = 0.f;
    for (int q = 0; q < n_parts; ++q) s += dpart[((size_t)q * H + h) * 32 + k];
    dtable[k * H + h] = s;
}

// ---------------------------------------------------------------- q_sample and the loss ----------------------------------------------------------------
// coef [B, 8] = sqrt_ac, sqrt_1mac, sqrt_recip_ac, sqrt_recipm1_ac, posterior_mean_coef1, coef2, min_log, max_log (per sample, fp32)
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const float* __restrict__ coef, float* __restrict__ xt,
                                size_t per, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / per;
        xt[i] = coef[b * 8] * x0[i] + coef[b * 8 + 1] * noise[i];
    }
}
TTTS_DEVICE float approx_cdf(float a) { return 0.5f * (1.f + tanhf(0.7978845608028654f * (a + 0.044715f * a * a * a))); }
TTTS_DEVICE float approx_cdf_grad(float a) {
    const float th = tanhf(0.7978845608028654f * (a + 0.044715f * a * a * a));
    return 0.5f * (1.f - th * th) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * a * a);
}
// one element of the variational-bound term and (optionally) its derivative with respect to the variance output v
TTTS_DEVICE float vb_term(float x0, float xt, float eps, float v, const float* c, int t0, float* dterm_dv) {
    const float min_log = c[6], max_log = c[7];
    const float frac = (v + 1.f) * 0.5f;
    const float logvar = frac * max_log + (1.f - frac) * min_log;
    const float px0 = fminf(fmaxf(c[2] * xt - c[3] * eps, -1.f), 1.f);
    const float mean = c[4] * px0 + c[5] * xt;
    const float dlv = 0.5f * (max_log - min_log);
    if (!t0) {
        const float tm = c[4] * x0 + c[5] * xt, dm = tm - mean;
        const float e1 = expf(min_log - logvar), e2 = expf(-logvar);
        if (dterm_dv) *dterm_dv = 0.5f * (1.f - e1 - dm * dm * e2) * dlv;
        return 0.5f * (-1.f + logvar - min_log + e1 + dm * dm * e2);
    }
    const float cx = x0 - mean, inv = expf(-0.5f * logvar);
    const float pin = inv * (cx + 1.f / 255.f), nin = inv * (cx - 1.f / 255.f);
    const float cp = approx_cdf(pin), cm = approx_cdf(nin);
    float val, dlogp;                                 // dlogp = d log_prob / d logvar (d pin / d logvar = -pin / 2)
    if (x0 < -0.999f) {
        val = fmaxf(cp, 1e-12f);
        dlogp = cp >= 1e-12f ? approx_cdf_grad(pin) * (-0.5f * pin) / val : 0.f;
    } else if (x0 > 0.999f) {
        val = fmaxf(1.f - cm, 1e-12f);
        dlogp = (1.f - cm) >= 1e-12f ? -approx_cdf_grad(nin) * (-0.5f * nin) / val : 0.f;
    } else {
        val = fmaxf(cp - cm, 1e-12f);
        dlogp = (cp - cm) >= 1e-12f ? (approx_cdf_grad(pin) * (-0.5f * pin) - approx_cdf_grad(nin) * (-0.5f * nin)) / val : 0.f;
    }
    if (dterm_dv) *dterm_dv = -dlogp * dlv;
    return -logf(val);
}
// partial sums: grid (chunks, B); part [B, chunks, 2] = (sum (noise - eps)^2, sum vb element terms) of the chunk
__global__ void __launch_bounds__(256) diff_loss_part_kernel(const float* __restrict__ out, const float* __restrict__ x0, const float* __restrict__ xt,
                                                             const float* __restrict__ noise, const float* __restrict__ coef, const int* __restrict__ t_is0,
                                                             float* __restrict__ part, int per) {
    __shared__ float red[8];
    const int b = blockIdx.y;
    const float* c = coef + (size_t)b * 8;
    const int t0 = t_is0[b];
    float s1 = 0.f, s2 = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < per; i += gridDim.x * 256) {
        const size_t k = (size_t)b * per + i;
        const float eps = out[(size_t)b * 2 * per + i], v = out[(size_t)b * 2 * per + per + i];
        const float d = noise[k] - eps;
        s1 = fmaf(d, d, s1);
        s2 += vb_term(x0[k], xt[k], eps, v, c, t0, nullptr);
    }
    s1 = block_sum256(s1, red);
    s2 = block_sum256(s2, red);
    if (threadIdx.x == 0) { part[((size_t)b * gridDim.x + blockIdx.x) * 2] = s1; part[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = s2; }
}
// terms [2, B] = (mse, vb) per sample, loss[0] = mean_b(mse + vb); one thread: B and chunks are tiny
__global__ void diff_loss_final_kernel(const float* __restrict__ part, float* __restrict__ terms, float* __restrict__ loss, int B, int chunks, int per) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float tot = 0.f;
    for (int b = 0; b < B; ++b) {
        float s1 = 0.f, s2 = 0.f;
        for (int q = 0; q < chunks; ++q) { s1 += part[((size_t)b * chunks + q) * 2]; s2 += part[((size_t)b * chunks + q) * 2 + 1]; }
        const float mse = s1 / (float)per, vb = s2 / (float)per / 0.6931471805599453f;
        terms[b] = mse; terms[B + b] = vb;
        tot += mse + vb;
    }
    loss[0] = tot / (float)B;
}
__global__ void diff_loss_bwd_kernel(const float* __restrict__ dL, const float* __restrict__ out, const float* __restrict__ x0, const float* __restrict__ xt,
                                     const float* __restrict__ noise, const float* __restrict__ coef, const int* __restrict__ t_is0,
                                     float* __restrict__ dout, int B, int per) {
    const size_t n = (size_t)B * per;
    const float g = dL[0] / (float)B / (float)per;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const size_t b = k / per, i = k - b * per;
        const float eps = out[b * 2 * per + i], v = out[b * 2 * per + per + i];
        float dv;
        vb_term(x0[k], xt[k], eps, v, coef + b * 8, t_is0[b], &dv);
        dout[b * 2 * per + i] = g * 2.f * (eps - noise[k]);
        dout[b * 2 * per + per + i] = g * dv / 0.6931471805599453f;
    }
}


// ---------------------------------------------------------------- layout conversion for the tensor-core convolutions ----------------------------------------------------------------
// The wide convolutions of the training tapes (AA_diffusion's 512-channel layers, the period discriminators' 512 / 1024-channel layers, the
// Generator's ResBlocks) are GEMMs; they run on ttts_gemm_bf16 (tcgen05) with split-bf16 operands (x = hi + lo, w = hi + lo,
// y ~ hi hi + hi lo + lo hi with fp32 accumulation: fp32-grade results, ttts_b200/vqvae/train_encoder.py).  The GEMM wants position-major
// rows, the tape keeps [B, C, T]: these two kernels convert.
//   cl_split : x [B, C, T] fp32 -> rows [hi(x[b, :, t]) | lo(x[b, :, t])] (2C bf16) at row row_off + b * rows_per_clip + t of a buffer whose other
//              rows stay ZERO (never written; the buffer is zero-initialised once): tap k of a convolution with stride s is the same buffer read
//              at rows s m + k (an A operand with row stride s), and no tap reads a neighbouring clip.  lrelu != 0: leaky_relu(0.1) first.
//   cl_unpack: D fp32 (row row_off + b * rows_per_clip + t, row pitch ld) -> y [B, C, T]
//              lrelu_x != nullptr: y = D * leaky_relu'(lrelu_x) (slope 0.1) -- the input gradient of a layer that applies the activation to
//              its input, folded into the conversion instead of one more pass over [B, C, T]
// 64 x 64 tiles through shared memory (16 KB in, 16 KB out per CTA), 128-byte warp requests on both sides: the [B, C, T] side along time, the
// position-major side along channels -- cl_split packs two neighbouring channels per lane into one 32-bit store.  The first version (32 x 32
// tiles, one 2-byte store per thread and part) ran at 2.4 (split) / 3.0 (unpack) TB/s and was 30 % of a routed 128-channel layer (r2ah).
// dil > 1: the clip is DE-INTERLEAVED on the position-major side -- sample t goes to sub-clip b dil + t mod dil, position t / dil (row
// row_off + (b dil + t mod dil) rows_per_clip + t / dil): on each residue class of the time index a dilated convolution is an ordinary one, so the
// dilated layers take the tap-concatenated route too; the [B, C, T] side is untouched and rows stay whole, so the kernels move the same bytes.
constexpr int CL_TILE = 64;
TTTS_DEVICE size_t cl_row(int b, int t, int rows_per_clip, int row_off, int dil) {
    if (dil <= 1) return (size_t)row_off + (size_t)b * rows_per_clip + t;
    const int q = t / dil, r = t - q * dil;
    return (size_t)row_off + ((size_t)b * dil + r) * rows_per_clip + q;
}
__global__ void __launch_bounds__(256) cl_split_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, int C, int T, int rows_per_clip,
                                                       int row_off, int lrelu, int dil) {
    __shared__ float tile[CL_TILE][CL_TILE + 1];
    const int t0 = blockIdx.x * CL_TILE, c0 = blockIdx.y * CL_TILE, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < CL_TILE; j += 8) {
        const int c = c0 + j;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = t0 + lane + 32 * h;
            float v = (c < C && t < T) ? x[((size_t)b * C + c) * T + t] : 0.f;
            if (lrelu) v = v > 0.f ? v : 0.1f * v;
            tile[j][lane + 32 * h] = v;
        }
    }
    __syncthreads();
    const bool packed = (C & 1) == 0;                                 // 32-bit stores need both halves of a row 4-byte aligned
    for (int i = warp; i < CL_TILE; i += 8) {
        const int t = t0 + i, c = c0 + 2 * lane;
        if (t >= T || c >= C) continue;
        const float v0 = tile[2 * lane][i], v1 = tile[2 * lane + 1][i];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
        const uint16_t uh0 = *reinterpret_cast<const uint16_t*>(&h0), uh1 = *reinterpret_cast<const uint16_t*>(&h1);
        const uint16_t ul0 = *reinterpret_cast<const uint16_t*>(&l0), ul1 = *reinterpret_cast<const uint16_t*>(&l1);
        uint16_t* row = out + cl_row(b, t, rows_per_clip, row_off, dil) * (2 * (size_t)C);
        if (packed) {                                                  // C even, c even: c + 1 < C
            *reinterpret_cast<uint32_t*>(row + c) = (uint32_t)uh0 | ((uint32_t)uh1 << 16);
            *reinterpret_cast<uint32_t*>(row + C + c) = (uint32_t)ul0 | ((uint32_t)ul1 << 16);
        } else {
            row[c] = uh0; row[C + c] = ul0;
            if (c + 1 < C) { row[c + 1] = uh1; row[C + c + 1] = ul1; }
        }
    }
}
__global__ void __launch_bounds__(256) cl_unpack_kernel(const float* __restrict__ D, float* __restrict__ y, int C, int T, int ld, int rows_per_clip,
                                                        int row_off, const float* __restrict__ lrelu_x, int dil) {
    __shared__ float tile[CL_TILE][CL_TILE + 1];
    const int t0 = blockIdx.x * CL_TILE, c0 = blockIdx.y * CL_TILE, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < CL_TILE; i += 8) {
        const int t = t0 + i;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + lane + 32 * h;
            tile[i][lane + 32 * h] = (t < T && c < C) ? D[cl_row(b, t, rows_per_clip, row_off, dil) * ld + c] : 0.f;
        }
    }
    __syncthreads();
    for (int j = warp; j < CL_TILE; j += 8) {
        const int c = c0 + j;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = t0 + lane + 32 * h;
            if (c < C && t < T) {
                const size_t o = ((size_t)b * C + c) * T + t;
                float v = tile[lane + 32 * h][j];
                if (lrelu_x != nullptr) v = lrelu_x[o] > 0.f ? v : 0.1f * v;
                y[o] = v;
            }
        }
    }
}

// Weights of a routed convolution for the tap-concatenated GEMM pair, in one pass: w [Cout, Cin, K] fp32 -> W1, W2 [R, K 2 Cr] bf16 with, per tap,
// W1 = [hi | hi], W2 = [lo | 0].  flip_t = 0: rows = output channels (R = Cout, Cr = Cin), tap order as stored (forward, weight gradient);
// flip_t = 1: the input-gradient form -- rows = input channels (R = Cin, Cr = Cout), taps reversed.  One thread per (row, column channel) walks the
// K taps of its weight (K contiguous floats).  Replaces ~8 torch kernels per call on a host-bound tape (r2al).
__global__ void __launch_bounds__(256) conv_w_concat_kernel(const float* __restrict__ w, uint16_t* __restrict__ W1, uint16_t* __restrict__ W2, int Cout,
                                                            int Cin, int K, int flip_t) {
    const int R = flip_t ? Cin : Cout, Cr = flip_t ? Cout : Cin;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)R * Cr) return;
    const int r = (int)(idx / Cr), c = (int)(idx - (size_t)r * Cr);
    const float* src = flip_t ? w + ((size_t)c * Cin + r) * K : w + ((size_t)r * Cin + c) * K;
    const size_t row = (size_t)r * K * 2 * Cr;
    for (int k = 0; k < K; ++k) {
        const float v = src[flip_t ? K - 1 - k : k];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const uint16_t uh = *reinterpret_cast<const uint16_t*>(&h), ul = *reinterpret_cast<const uint16_t*>(&l);
        const size_t o = row + (size_t)k * 2 * Cr + c;
        W1[o] = uh; W1[o + Cr] = uh;
        W2[o] = ul; W2[o + Cr] = 0;
    }
}

static inline unsigned df_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (unsigned)(b > cap ? cap : (b ? b : 1));
}

static int attn_bias_setup(AttnBiasParams& p, int B, int C, int T, int H, size_t& smem, int kind) {
    TTTS_CHECK_ARG(B >= 1 && B <= 65535 && H >= 1 && H <= 65535 && C >= 1 && C % H == 0 && T >= 1, "attn_bias: bad shape");
    const int ch = C / H;
    TTTS_CHECK_ARG(ch == 8 || ch == 16 || ch == 32 || ch == 64, "attn_bias: head width %d not in {8, 16, 32, 64}", ch);
    p.C = C; p.T = T; p.H = H; p.ch = ch;
    // kind 0 forward: Q K V + P + bias; 1 dq: Q K V dO + S + bias + diag + lse/delta; 2 dkv: Q K V dO + S P + bias + lse/delta
    const size_t tiles = kind == 0 ? 3 : 4, sq = kind == 2 ? 2 : 1;
    smem = (tiles * ch * PITCH + sq * AT * PITCH + (kind == 1 ? 4 : 2) * (size_t)T + 2 * AT) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 220 * 1024, "attn_bias: T = %d does not fit the shared-memory bias table", T);
    return TTTS_OK;
}
template <typename Kern>
static int set_smem(Kern k, size_t smem) {
#ifndef TTTS_HOST_EMU
    if (smem > 48 * 1024) TTTS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
    return TTTS_OK;
}

}  // namespace ttts

using namespace ttts;

/* y = act(GroupNorm(x) (1 + scale) + shift) ; x, y [B,C,T] ; gamma, beta [C] ; scale, shift [B,C] or both NULL ; stats [B,G,2] out */
extern "C" int ttts_groupnorm(const float* x, const float* gamma, const float* beta, const float* scale, const float* shift, float* y, float* stats,
                              int32_t B, int32_t C, int32_t T, int32_t G, int32_t silu, void* stream) {
    TTTS_CHECK_ARG(x && gamma && beta && y && stats && B >= 1 && B <= 65535 && C >= 1 && T >= 1 && G >= 1 && C % G == 0 && C / G <= 64,
                   "groupnorm: bad args");
    TTTS_CHECK_ARG((scale == nullptr) == (shift == nullptr), "groupnorm: scale and shift go together");
    TTTS_CUDA(launch_plain(gn_fwd_kernel, dim3(G, B), dim3(256), 0, (cudaStream_t)stream, x, gamma, beta, scale, shift, y, stats, C, T, G, silu));
    TTTS_LAUNCH_CHECK("gn_fwd");
    return TTTS_OK;
}
/* dx [B,C,T], dgamma / dbeta [C], dscale / dshift [B,C] (NULL without modulation) written ; scratch: B*C*2 floats */
extern "C" int ttts_groupnorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* beta, const float* scale,
                                  const float* shift, float* dx, float* dgamma, float* dbeta, float* dscale, float* dshift, float* scratch,
                                  int32_t B, int32_t C, int32_t T, int32_t G, int32_t silu, void* stream) {
    TTTS_CHECK_ARG(dy && x && stats && gamma && beta && dx && dgamma && dbeta && scratch && B >= 1 && B <= 65535 && C >= 1 && T >= 1 && G >= 1 &&
                   C % G == 0 && C / G <= 64, "groupnorm backward: bad args");
    TTTS_CHECK_ARG((scale == nullptr) == (shift == nullptr) && (scale == nullptr) == (dscale == nullptr) && (dscale == nullptr) == (dshift == nullptr),
                   "groupnorm backward: scale / shift / dscale / dshift go together");
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CUDA(launch_plain(gn_bwd_kernel, dim3(G, B), dim3(256), 0, st, dy, x, stats, gamma, beta, scale, shift, dx, scratch, dscale, dshift, C, T, G, silu));
    TTTS_LAUNCH_CHECK("gn_bwd");
    TTTS_CUDA(launch_plain(gn_param_kernel, dim3((C + 127) / 128), dim3(128), 0, st, (const float*)scratch, dgamma, dbeta, B, C));
    TTTS_LAUNCH_CHECK("gn_param");
    return TTTS_OK;
}
extern "C" int ttts_silu(const float* x, const float* dy, float* out, int64_t n, int32_t backward, void* stream) {
    TTTS_CHECK_ARG(x && out && n >= 1 && (!backward || dy), "silu: bad args");
    TTTS_CUDA(launch_plain(silu_kernel, dim3(df_blocks((size_t)n)), dim3(256), 0, (cudaStream_t)stream, x, dy, out, (size_t)n, backward));
    TTTS_LAUNCH_CHECK("silu");
    return TTTS_OK;
}

/* out [B,C,T], lse [B,H,T] = attention(qkv [B,3C,T]) with the relative-position bias table [32,H] ; diag int32 [2T-1] */
extern "C" int ttts_attn_bias(const float* qkv, const float* table, const int32_t* diag, float* out, float* lse, int32_t B, int32_t C, int32_t T,
                              int32_t H, void* stream) {
    AttnBiasParams p = {};
    p.qkv = qkv; p.table = table; p.diag = diag; p.o = out; p.lse = lse;
    size_t smem;
    TTTS_RUN(attn_bias_setup(p, B, C, T, H, smem, 0));
    TTTS_CHECK_ARG(qkv && table && diag && out && lse, "attn_bias: null pointer");
    const dim3 grid((T + AT - 1) / AT, H, B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (p.ch) {
        case 64: TTTS_RUN(set_smem(attn_bias_fwd_kernel<4>, smem)); TTTS_CUDA(launch_plain(attn_bias_fwd_kernel<4>, grid, dim3(256), smem, st, p)); break;
        case 32: TTTS_RUN(set_smem(attn_bias_fwd_kernel<2>, smem)); TTTS_CUDA(launch_plain(attn_bias_fwd_kernel<2>, grid, dim3(256), smem, st, p)); break;
        default: TTTS_RUN(set_smem(attn_bias_fwd_kernel<1>, smem)); TTTS_CUDA(launch_plain(attn_bias_fwd_kernel<1>, grid, dim3(256), smem, st, p)); break;
    }
    TTTS_LAUNCH_CHECK("attn_bias_fwd");
    return TTTS_OK;
}
/* floats of scratch ttts_attn_bias_bwd needs: delta [B,H,T] + per-CTA bucket sums */
extern "C" int64_t ttts_attn_bias_bwd_scratch_floats(int32_t B, int32_t T, int32_t H) {
    return (int64_t)B * H * T + (int64_t)B * ((T + AT - 1) / AT) * H * 32;
}
/* dqkv [B,3C,T], dtable [32,H] written */
extern "C" int ttts_attn_bias_bwd(const float* dout, const float* qkv, const float* out, const float* lse, const float* table, const int32_t* diag,
                                  float* dqkv, float* dtable, float* scratch, int32_t B, int32_t C, int32_t T, int32_t H, void* stream) {
    AttnBiasParams p = {};
    p.qkv = qkv; p.table = table; p.diag = diag; p.dout = dout; p.lse_in = lse; p.dqkv = dqkv;
    size_t smem_q, smem_kv;
    TTTS_RUN(attn_bias_setup(p, B, C, T, H, smem_q, 1));
    TTTS_RUN(attn_bias_setup(p, B, C, T, H, smem_kv, 2));
    TTTS_CHECK_ARG(dout && qkv && out && lse && table && diag && dqkv && dtable && scratch, "attn_bias backward: null pointer");
    float* delta = scratch;
    p.delta = delta;
    p.dpart = scratch + (size_t)B * H * T;
    const int nq = (T + AT - 1) / AT;
    const dim3 grid(nq, H, B);
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CUDA(launch_plain(attn_bias_delta_kernel, dim3(df_blocks((size_t)B * H * T)), dim3(256), 0, st, dout, out, delta, C, T, H, (size_t)B * H * T));
    TTTS_LAUNCH_CHECK("attn_bias_delta");
    switch (p.ch) {
        case 64:
            TTTS_RUN(set_smem(attn_bias_bwd_dq_kernel<4>, smem_q)); TTTS_CUDA(launch_plain(attn_bias_bwd_dq_kernel<4>, grid, dim3(256), smem_q, st, p));
            TTTS_RUN(set_smem(attn_bias_bwd_dkv_kernel<4>, smem_kv)); TTTS_CUDA(launch_plain(attn_bias_bwd_dkv_kernel<4>, grid, dim3(256), smem_kv, st, p));
            break;
        case 32:
            TTTS_RUN(set_smem(attn_bias_bwd_dq_kernel<2>, smem_q)); TTTS_CUDA(launch_plain(attn_bias_bwd_dq_kernel<2>, grid, dim3(256), smem_q, st, p));
            TTTS_RUN(set_smem(attn_bias_bwd_dkv_kernel<2>, smem_kv)); TTTS_CUDA(launch_plain(attn_bias_bwd_dkv_kernel<2>, grid, dim3(256), smem_kv, st, p));
            break;
        default:
            TTTS_RUN(set_smem(attn_bias_bwd_dq_kernel<1>, smem_q)); TTTS_CUDA(launch_plain(attn_bias_bwd_dq_kernel<1>, grid, dim3(256), smem_q, st, p));
            TTTS_RUN(set_smem(attn_bias_bwd_dkv_kernel<1>, smem_kv)); TTTS_CUDA(launch_plain(attn_bias_bwd_dkv_kernel<1>, grid, dim3(256), smem_kv, st, p));
            break;
    }
    TTTS_LAUNCH_CHECK("attn_bias_bwd");
    TTTS_CUDA(launch_plain(attn_bias_dtable_kernel, dim3((H * 32 + 127) / 128), dim3(128), 0, st, (const float*)p.dpart, dtable, B * nq, H));
    TTTS_LAUNCH_CHECK("attn_bias_dtable");
    return TTTS_OK;
}

/* x_t = coef[b,0] x_start + coef[b,1] noise ; tensors [B, per] */
extern "C" int ttts_diff_q_sample(const float* x_start, const float* noise, const float* coef, float* x_t, int32_t B, int64_t per, void* stream) {
    TTTS_CHECK_ARG(x_start && noise && coef && x_t && B >= 1 && per >= 1, "q_sample: bad args");
    const size_t n = (size_t)B * per;
    TTTS_CUDA(launch_plain(q_sample_kernel, dim3(df_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x_start, noise, coef, x_t, (size_t)per, n));
    TTTS_LAUNCH_CHECK("q_sample");
    return TTTS_OK;
}
/* model_out [B, 2 Cn, T] (eps | variance values), x_start / x_t / noise [B, Cn, T], coef [B,8], t_is0 int32 [B] ->
 * terms [2,B] = (mse, vb), loss [1] = mean_b(mse + vb) ; scratch: B * TTTS_DIFF_LOSS_CHUNKS * 2 floats */
#define TTTS_DIFF_LOSS_CHUNKS 32
extern "C" int ttts_diff_loss(const float* model_out, const float* x_start, const float* x_t, const float* noise, const float* coef, const int32_t* t_is0,
                              float* terms, float* loss, float* scratch, int32_t B, int32_t Cn, int32_t T, void* stream) {
    TTTS_CHECK_ARG(model_out && x_start && x_t && noise && coef && t_is0 && terms && loss && scratch && B >= 1 && B <= 65535 && Cn >= 1 && T >= 1,
                   "diff_loss: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    const int per = Cn * T;
    TTTS_CUDA(launch_plain(diff_loss_part_kernel, dim3(TTTS_DIFF_LOSS_CHUNKS, B), dim3(256), 0, st, model_out, x_start, x_t, noise, coef, t_is0, scratch, per));
    TTTS_LAUNCH_CHECK("diff_loss_part");
    TTTS_CUDA(launch_plain(diff_loss_final_kernel, dim3(1), dim3(32), 0, st, (const float*)scratch, terms, loss, B, TTTS_DIFF_LOSS_CHUNKS, per));
    TTTS_LAUNCH_CHECK("diff_loss_final");
    return TTTS_OK;
}
/* d model_out [B, 2 Cn, T] for a loss gradient dL [1] */
extern "C" int ttts_diff_loss_bwd(const float* dL, const float* model_out, const float* x_start, const float* x_t, const float* noise, const float* coef,
                                  const int32_t* t_is0, float* dout, int32_t B, int32_t Cn, int32_t T, void* stream) {
    TTTS_CHECK_ARG(dL && model_out && x_start && x_t && noise && coef && t_is0 && dout && B >= 1 && Cn >= 1 && T >= 1, "diff_loss backward: bad args");
    const int per = Cn * T;
    TTTS_CUDA(launch_plain(diff_loss_bwd_kernel, dim3(df_blocks((size_t)B * per)), dim3(256), 0, (cudaStream_t)stream, dL, model_out, x_start, x_t, noise,
                           coef, t_is0, dout, B, per));
    TTTS_LAUNCH_CHECK("diff_loss_bwd");
    return TTTS_OK;
}

/* x [B,C,T] fp32 -> split-bf16 position-major rows [hi | lo] of a zero-initialised [rows, 2C] bf16 buffer: row row_off + b rows_per_clip + t
 * (dil > 1: row_off + (b dil + t mod dil) rows_per_clip + t / dil, the de-interleaved order in which a dilated convolution is an ordinary one) */
extern "C" int ttts_cl_split(const float* x, void* out_bf16, int32_t B, int32_t C, int32_t T, int32_t rows_per_clip, int32_t row_off, int32_t lrelu,
                             int32_t dil, void* stream) {
    TTTS_CHECK_ARG(x && out_bf16 && B >= 1 && B <= 65535 && C >= 1 && T >= 1 && (C + CL_TILE - 1) / CL_TILE <= 65535 && dil >= 1 &&
                   rows_per_clip >= (T + dil - 1) / dil && row_off >= 0, "cl_split: bad args");
    TTTS_CHECK_ARG(((uintptr_t)out_bf16 & 3) == 0, "cl_split: output not 4-byte aligned");
    TTTS_CUDA(launch_plain(cl_split_kernel, dim3((T + CL_TILE - 1) / CL_TILE, (C + CL_TILE - 1) / CL_TILE, B), dim3(256), 0, (cudaStream_t)stream, x,
                           (uint16_t*)out_bf16, C, T, rows_per_clip, row_off, lrelu, dil));
    TTTS_LAUNCH_CHECK("cl_split");
    return TTTS_OK;
}
/* D fp32 position-major (row row_off + b rows_per_clip + t, pitch ld) -> y [B,C,T]; lrelu_x (may be NULL): y *= leaky_relu'(lrelu_x), slope 0.1 */
extern "C" int ttts_cl_unpack(const float* D, float* y, int32_t B, int32_t C, int32_t T, int32_t ld, int32_t rows_per_clip, int32_t row_off,
                              const float* lrelu_x, int32_t dil, void* stream) {
    TTTS_CHECK_ARG(D && y && B >= 1 && B <= 65535 && C >= 1 && T >= 1 && ld >= C && (C + CL_TILE - 1) / CL_TILE <= 65535 && dil >= 1 &&
                   rows_per_clip >= (T + dil - 1) / dil && row_off >= 0, "cl_unpack: bad args");
    TTTS_CUDA(launch_plain(cl_unpack_kernel, dim3((T + CL_TILE - 1) / CL_TILE, (C + CL_TILE - 1) / CL_TILE, B), dim3(256), 0, (cudaStream_t)stream, D, y, C, T,
                           ld, rows_per_clip, row_off, lrelu_x, dil));
    TTTS_LAUNCH_CHECK("cl_unpack");
    return TTTS_OK;
}
/* w [Cout,Cin,K] fp32 -> the two bf16 B operands of the tap-concatenated GEMM pair (see conv_w_concat_kernel) */
extern "C" int ttts_conv_w_concat(const float* w, void* W1, void* W2, int32_t Cout, int32_t Cin, int32_t K, int32_t flip_transpose, void* stream) {
    TTTS_CHECK_ARG(w && W1 && W2 && Cout >= 1 && Cin >= 1 && K >= 1 && (size_t)Cout * Cin <= ((size_t)1 << 31), "conv_w_concat: bad args");
    const size_t n = (size_t)Cout * Cin;
    TTTS_CUDA(launch_plain(conv_w_concat_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, w, (uint16_t*)W1, (uint16_t*)W2,
                           Cout, Cin, K, flip_transpose));
    TTTS_LAUNCH_CHECK("conv_w_concat");
    return TTTS_OK;
}
