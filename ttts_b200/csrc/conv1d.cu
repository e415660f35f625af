// fp32 Conv1d family for the VQ-VAE encode front end (PosteriorAudioEncoder / MelStyleEncoder / WN):
//   y[b, co, t] = bias[co] + sum_{ci, k} w[co, ci, k] * pre(x[b, ci, t*stride + k*dil - pad])
// with the elementwise work of the reference's blocks fused in:
//   pre   : identity | leaky_relu(0.1)                        (ResBlock1: modules.py:302-309)
//   post  : + residual, * out_scale (accumulating the mean of 3 ResBlock1 branches, vq2.py:723-729), * time mask,
//           GLU (x1 * sigmoid(x2), modules.py:560-566), Mish (modules.py Mish), gated tanh*sigmoid with a per-batch
//           conditioning vector (WN: commons.fused_add_tanh_sigmoid_multiply, modules.py:196-203)
// The reference runs these layers in fp32 (fp16_run=false, vqvae/config.json:21); to keep the downstream "bit-exact VQ
// indices" contract meaningful this stays on the FP32 FMA pipe.  Channel counts are tiny (16..192, 1025 for the 1x1 `pre`),
// so a register-tiled direct convolution (implicit GEMM on CUDA cores) is the right tool: CTA tile = 64 output channels x
// 64 time steps, input window and weight slab staged through shared memory per 16-input-channel chunk, 4x4 outputs/thread.
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#include "conv_params.h"

namespace ttts {

// conv1d_split.cu (off unless TTTS_CONV_SPLIT=1 or forced by ttts_conv1d_f32_split): split-reduction form of the <32> pipelined kernel
int conv1d_split_try(const ConvParams& p, dim3 grid, int force_groups, cudaStream_t st);

constexpr int CV_CO = 64, CV_T = 64, CV_CI = 16, CV_THREADS = 256;

// For the gated posts (GLU / WN) a CTA computes BOTH halves of 32 gate channels: output-channel tile of 64 = 32 "a" + 32 "b".
__global__ void __launch_bounds__(CV_THREADS) conv1d_f32_kernel(const ConvParams p) {
    extern __shared__ float cv_smem[];
    const int gated = (p.post == 1 || p.post == 3);
    const int Chalf = p.Cout >> 1;
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * CV_T;
    const int co0 = blockIdx.y * (gated ? CV_CO / 2 : CV_CO);
    const int win = (CV_T - 1) * p.stride + (p.K - 1) * p.dil + 1;       // input window length per channel
    float* sx = cv_smem;                                                 // [CV_CI][win]
    float* sw = cv_smem + CV_CI * win;                                   // [CV_CI][K][CV_CO]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;                              // tx: 4 time steps each, ty: 4 channels each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int in0 = t0 * p.stride - p.pad;

    for (int c0 = 0; c0 < p.Cin; c0 += CV_CI) {
        __syncthreads();
        for (int i = tid; i < CV_CI * win; i += CV_THREADS) {
            const int ci = i / win, u = i - ci * win;
            const int c = c0 + ci, ti = in0 + u;
            float v = 0.f;
            if (c < p.Cin && ti >= 0 && ti < p.Tin) {
                v = p.x[((size_t)b * p.Cin + c) * p.Tin + ti];
                if (p.pre_lrelu) v = v > 0.f ? v : 0.1f * v;
            }
            sx[i] = v;
        }
        for (int i = tid; i < CV_CI * p.K * CV_CO; i += CV_THREADS) {
            const int col = i % CV_CO, rk = i / CV_CO;
            const int k = rk % p.K, ci = rk / p.K;
            int co;
            if (gated) co = (col < CV_CO / 2) ? co0 + col : Chalf + co0 + (col - CV_CO / 2);
            else co = co0 + col;
            const int c = c0 + ci;
            const bool ok = gated ? ((col < CV_CO / 2 ? co0 + col : co0 + col - CV_CO / 2) < Chalf) : (co < p.Cout);
            sw[i] = (ok && c < p.Cin) ? __ldg(p.w + ((size_t)co * p.Cin + c) * p.K + k) : 0.f;
        }
        __syncthreads();
        const int cimax = min(CV_CI, p.Cin - c0);
        for (int ci = 0; ci < cimax; ++ci) {
            for (int k = 0; k < p.K; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(sw + (ci * p.K + k) * CV_CO + ty * 4);
                const float* xr = sx + ci * win + k * p.dil + (tx * 4) * p.stride;
                const float x0 = xr[0], x1 = xr[p.stride], x2 = xr[2 * p.stride], x3 = xr[3 * p.stride];
                const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(wa[i], x0, acc[i][0]);
                    acc[i][1] = fmaf(wa[i], x1, acc[i][1]);
                    acc[i][2] = fmaf(wa[i], x2, acc[i][2]);
                    acc[i][3] = fmaf(wa[i], x3, acc[i][3]);
                }
            }
        }
    }

    // ---------------- epilogue ----------------
    if (!gated) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + ty * 4 + i;
            if (co >= p.Cout) continue;
            const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = t0 + tx * 4 + j;
                if (t >= p.Tout) continue;
                float v = acc[i][j] + bv;
                if (p.post == 2) v = mish_f(v);
                const size_t o = ((size_t)b * p.Cout + co) * p.Tout + t;
                if (p.resid) v += p.resid[o];
                v *= p.out_scale;
                if (p.mask) v *= p.mask[(size_t)b * p.Tout + t];
                p.y[o] = p.accumulate ? p.y[o] + v : v;
            }
        }
    } else {
        // thread rows ty*4..+3 of the 64-wide tile: rows < 32 are "a" channels, rows >= 32 the matching "b" channels -> exchange via smem
        __syncthreads();
        float* sg = cv_smem;                               // [64][64] accumulators
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sg[(ty * 4 + i) * CV_T + tx * 4 + j] = acc[i][j];
        __syncthreads();
        for (int i = tid; i < (CV_CO / 2) * CV_T; i += CV_THREADS) {
            const int cl = i / CV_T, tl = i - cl * CV_T;
            const int c = co0 + cl, t = t0 + tl;
            if (c >= Chalf || t >= p.Tout) continue;
            float a = sg[cl * CV_T + tl] + (p.bias ? __ldg(p.bias + c) : 0.f);
            float g = sg[(cl + CV_CO / 2) * CV_T + tl] + (p.bias ? __ldg(p.bias + Chalf + c) : 0.f);
            float v;
            if (p.post == 1) {
                v = a * (1.f / (1.f + expf(-g)));                                         // GLU
            } else {
                if (p.cond) { a += p.cond[(size_t)b * p.cond_ld + c]; g += p.cond[(size_t)b * p.cond_ld + Chalf + c]; }
                v = tanhf(a) * (1.f / (1.f + expf(-g)));                                  // WN gate
            }
            const size_t o = ((size_t)b * Chalf + c) * p.Tout + t;
            if (p.resid) v += p.resid[o];
            v *= p.out_scale;
            if (p.mask) v *= p.mask[(size_t)b * p.Tout + t];
            p.y[o] = p.accumulate ? p.y[o] + v : v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Implicit-GEMM form (default):  Y[co, p] = sum_r W[co, r] * Xcol[r, p],  r = ci*K + k,  p = b*Tout + t  (batch and time FLATTENED,
// so T = 36 frames or T = 1 pack into full 64-position tiles, and any K / stride / dilation is a runtime property of the loader, not of
// the inner loop).  CTA tile = CO_T (64 or 32) output channels x 64 positions, 16 r-rows per chunk, double-buffered shared memory with
// the next chunk's global loads in flight during the FMAs, 4x4 outputs per thread.  The accumulation order over r is the same as in the
// window kernel above, so both produce bit-identical results.
// Why: the encoder launch list (profiles/r1c_launches_vqenc_summary.txt) showed the window kernel at ~5.6 TFLOP/s: 64x64 tiles half
// empty for Cout = 32 or T = 36, tiny grids, runtime-K inner loops.
// ------------------------------------------------------------------------------------------------------------

template <int CO_T>
__global__ void __launch_bounds__(CO_T * 4) conv1d_igemm_kernel(const ConvParams p) {
    constexpr int NT = CO_T * 4;                 // threads: (CO_T/4) x 16
    constexpr int LDA = CO_T + 4, LDB = IG_P + 4;
    constexpr int NB = IG_R * IG_P / NT;         // B-tile elements per thread (4 or 8)
    __shared__ __align__(16) float sA[2][IG_R][LDA];
    __shared__ __align__(16) float sB[2][IG_R][LDB];
    const int gated = (p.post == 1 || p.post == 3);
    const int Chalf = p.Cout >> 1;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int p0 = blockIdx.x * IG_P;
    const int co0 = blockIdx.y * (gated ? CO_T / 2 : CO_T);
    const int R = p.Cin * p.K;
    const int Ptot = p.B * p.Tout;
    const int nchunks = (R + IG_R - 1) / IG_R;

    // ---- B loader: this thread always loads column cb (one output position), rows rb0 + (NT/64)*i
    const int cb = tid & 63, rb0 = tid >> 6;
    const int posb = p0 + cb;
    const bool pos_ok = posb < Ptot;
    const int bb = pos_ok ? posb / p.Tout : 0;
    const int tb = pos_ok ? posb - bb * p.Tout : 0;
    const float* xb = p.x + (size_t)bb * p.Cin * p.Tin;
    const int ti0 = tb * p.stride - p.pad;
    // ---- A loader: row (output channel) ca, r-columns ra4..ra4+3
    const int ca = tid >> 2, ra4 = (tid & 3) * 4;
    int coa; bool coa_ok;
    if (gated) { const int cl = ca < CO_T / 2 ? ca : ca - CO_T / 2; coa_ok = (co0 + cl) < Chalf; coa = (ca < CO_T / 2 ? 0 : Chalf) + co0 + cl; }
    else { coa = co0 + ca; coa_ok = coa < p.Cout; }
    const float* wa = p.w + (size_t)coa * R;

    float ra[4], rbv[NB];
    auto gload = [&](int chunk) {
        const int r0 = chunk * IG_R;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int rr = r0 + ra4 + i; ra[i] = (coa_ok && rr < R) ? __ldg(wa + rr) : 0.f; }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int rr = r0 + rb0 + (NT / 64) * i;
            const int ci = rr / p.K, k = rr - ci * p.K;
            const int ti = ti0 + k * p.dil;
            float v = 0.f;
            if (pos_ok && rr < R && ti >= 0 && ti < p.Tin) {
                v = __ldg(xb + (size_t)ci * p.Tin + ti);
                if (p.pre_lrelu) v = v > 0.f ? v : 0.1f * v;
            }
            rbv[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sA[buf][ra4 + i][ca] = ra[i];
#pragma unroll
        for (int i = 0; i < NB; ++i) sB[buf][rb0 + (NT / 64) * i][cb] = rbv[i];
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    gload(0);
    sstore(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) gload(c + 1);
#pragma unroll
        for (int r = 0; r < IG_R; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(&sA[buf][r][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&sB[buf][r][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        if (c + 1 < nchunks) sstore(buf ^ 1);
        __syncthreads();
    }

    // ---------------- epilogue ----------------
    if (!gated) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = p0 + tx * 4 + j;
            if (pos >= Ptot) continue;
            const int b = pos / p.Tout, t = pos - b * p.Tout;
            const float mk = p.mask ? p.mask[(size_t)b * p.Tout + t] : 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = co0 + ty * 4 + i;
                if (co >= p.Cout) continue;
                float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
                if (p.post == 2) v = mish_f(v);
                const size_t o = ((size_t)b * p.Cout + co) * p.Tout + t;
                if (p.resid) v += p.resid[o];
                v *= p.out_scale;
                if (p.mask) v *= mk;
                p.y[o] = p.accumulate ? p.y[o] + v : v;
            }
        }
    } else {
        // rows < CO_T/2 of the tile are "a" channels, rows >= CO_T/2 the matching "b" channels -> exchange through shared memory
        __shared__ float sgate[CO_T][IG_P + 1];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sgate[ty * 4 + i][tx * 4 + j] = acc[i][j];
        __syncthreads();
        for (int i = tid; i < (CO_T / 2) * IG_P; i += NT) {
            const int cl = i / IG_P, pl = i - cl * IG_P;
            const int c = co0 + cl, pos = p0 + pl;
            if (c >= Chalf || pos >= Ptot) continue;
            const int b = pos / p.Tout, t = pos - b * p.Tout;
            float a = sgate[cl][pl] + (p.bias ? __ldg(p.bias + c) : 0.f);
            float g = sgate[cl + CO_T / 2][pl] + (p.bias ? __ldg(p.bias + Chalf + c) : 0.f);
            float v;
            if (p.post == 1) {
                v = a * (1.f / (1.f + expf(-g)));                                         // GLU
            } else {
                if (p.cond) { a += p.cond[(size_t)b * p.cond_ld + c]; g += p.cond[(size_t)b * p.cond_ld + Chalf + c]; }
                v = tanhf(a) * (1.f / (1.f + expf(-g)));                                  // WN gate
            }
            const size_t o = ((size_t)b * Chalf + c) * p.Tout + t;
            if (p.resid) v += p.resid[o];
            v *= p.out_scale;
            if (p.mask) v *= p.mask[(size_t)b * p.Tout + t];
            p.y[o] = p.accumulate ? p.y[o] + v : v;
        }
    }
}

// Pipelined form of the implicit GEMM (default; TTTS_CONV_PIPE=0 selects the double-buffered kernel above): the encoder's layers are
// small (Cout <= 192, 2 304 positions for the 16 WN layers -> 216 CTAs of 4 warps), so each CTA's chunk loop ran at the latency of one
// L2 round trip per 16-row chunk (~1.7 us per chunk, profiles/r1d_launches_vqenc.csv: 104 us for 0.85 GFLOP).  Here the im2col gather
// is IG_STAGES - 1 chunks ahead through 4-byte cp.async (zero fill for padding), the (ci, k) decomposition of the row index is advanced
// incrementally instead of divided out per element, leaky-ReLU moves to the shared-memory read.  Same accumulation order: bit-identical.

template <int CO_T>
__global__ void __launch_bounds__(CO_T * 4) conv1d_igemm_pipe_kernel(const ConvParams p) {
    constexpr int NT = CO_T * 4;                 // threads: (CO_T/4) x 16
    constexpr int LDA = CO_T + 4, LDB = IG_P + 4;
    constexpr int NB = IG_R * IG_P / NT;         // B-tile elements per thread (4 or 8)
    constexpr int S = IG_STAGES;
    constexpr int A_ST = IG_R * LDA, B_ST = IG_R * LDB;                         // floats per stage
    constexpr int GATE_F = CO_T * (IG_P + 1);                                   // gated epilogue exchange buffer aliases the ring
    constexpr int SMEM_F = S * (A_ST + B_ST) > GATE_F ? S * (A_ST + B_ST) : GATE_F;
    __shared__ __align__(16) float ig_smem[SMEM_F];
    float* const sA = ig_smem;                   // [S][IG_R][LDA]
    float* const sB = ig_smem + S * A_ST;        // [S][IG_R][LDB]
    const int gated = (p.post == 1 || p.post == 3);
    const int Chalf = p.Cout >> 1;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int p0 = blockIdx.x * IG_P;
    const int co0 = blockIdx.y * (gated ? CO_T / 2 : CO_T);
    const int R = p.Cin * p.K;
    const int Ptot = p.B * p.Tout;
    const int nchunks = (R + IG_R - 1) / IG_R;

    // ---- B loader: this thread always loads column cb (one output position), rows rb0 + (NT/64)*i
    const int cb = tid & 63, rb0 = tid >> 6;
    const int posb = p0 + cb;
    const bool pos_ok = posb < Ptot;
    const int bb = pos_ok ? posb / p.Tout : 0;
    const int tb = pos_ok ? posb - bb * p.Tout : 0;
    const float* xb = p.x + (size_t)bb * p.Cin * p.Tin;
    const int ti0 = tb * p.stride - p.pad;
    // ---- A loader: row (output channel) ca, r-columns ra4..ra4+3
    const int ca = tid >> 2, ra4 = (tid & 3) * 4;
    int coa; bool coa_ok;
    if (gated) { const int cl = ca < CO_T / 2 ? ca : ca - CO_T / 2; coa_ok = (co0 + cl) < Chalf; coa = (ca < CO_T / 2 ? 0 : Chalf) + co0 + cl; }
    else { coa = co0 + ca; coa_ok = coa < p.Cout; }
    const float* wa = p.w + (size_t)coa * R;

    // (ci, k) of each of this thread's B elements, advanced by IG_R rows per issued chunk (no division in the loop)
    int bci[NB], bk[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { const int rr = rb0 + (NT / 64) * i; bci[i] = rr / p.K; bk[i] = rr - bci[i] * p.K; }
    const int c16 = IG_R / p.K, k16 = IG_R - c16 * p.K;
    // chunks are issued in order: chunk -> stage chunk % S, 4-byte cp.async with zero fill for padding / tails
    auto issue = [&](int chunk) {
        const int r0 = chunk * IG_R;
        float* a = sA + (chunk % S) * A_ST;
        float* b = sB + (chunk % S) * B_ST;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = r0 + ra4 + i;
            const bool ok = coa_ok && rr < R;
            cp_async4(&a[(ra4 + i) * LDA + ca], ok ? wa + rr : p.w, ok);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int rr = r0 + rb0 + (NT / 64) * i;
            const int ti = ti0 + bk[i] * p.dil;
            const bool ok = pos_ok && rr < R && ti >= 0 && ti < p.Tin;
            cp_async4(&b[(rb0 + (NT / 64) * i) * LDB + cb], ok ? xb + (size_t)bci[i] * p.Tin + ti : p.x, ok);
            bk[i] += k16; bci[i] += c16;
            if (bk[i] >= p.K) { bk[i] -= p.K; ++bci[i]; }
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < S - 1; ++s) { if (s < nchunks) issue(s); cp_async_commit(); }
    const bool lrelu = p.pre_lrelu != 0;
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<S - 2>();                  // chunk c has landed (S - 2 younger groups may still be in flight)
        if (lrelu) {
            // leaky ReLU once per element, by the thread that copied it (its own cp.async data is visible to it after the wait): in the
            // FMA loop it cost 12 instructions per 16 FMAs (r1n profile: FSETP + FMUL = 15 % of all instructions, FFMA 28 %)
            float* bw = sB + (c % S) * B_ST;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                float* e = &bw[(rb0 + (NT / 64) * i) * LDB + cb];
                const float v = *e;
                *e = v > 0.f ? v : 0.1f * v;
            }
        }
        __syncthreads();                         // ... for every thread, and everybody is done computing on stage (c - 1) % S
        if (c + S - 1 < nchunks) issue(c + S - 1);
        cp_async_commit();
        const float* a = sA + (c % S) * A_ST;
        const float* b = sB + (c % S) * B_ST;
#pragma unroll
        for (int r = 0; r < IG_R; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(&a[r * LDA + ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&b[r * LDB + tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w};
            const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
    }

    // ---------------- epilogue ----------------
    if (!gated) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = p0 + tx * 4 + j;
            if (pos >= Ptot) continue;
            const int b = pos / p.Tout, t = pos - b * p.Tout;
            const float mk = p.mask ? p.mask[(size_t)b * p.Tout + t] : 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = co0 + ty * 4 + i;
                if (co >= p.Cout) continue;
                float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
                if (p.post == 2) v = mish_f(v);
                const size_t o = ((size_t)b * p.Cout + co) * p.Tout + t;
                if (p.resid) v += p.resid[o];
                v *= p.out_scale;
                if (p.mask) v *= mk;
                p.y[o] = p.accumulate ? p.y[o] + v : v;
            }
        }
    } else {
        // rows < CO_T/2 of the tile are "a" channels, rows >= CO_T/2 the matching "b" channels -> exchange through shared memory
        float (*sgate)[IG_P + 1] = reinterpret_cast<float (*)[IG_P + 1]>(ig_smem);
        __syncthreads();                         // the ring is dead: every thread has left the main loop
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sgate[ty * 4 + i][tx * 4 + j] = acc[i][j];
        __syncthreads();
        for (int i = tid; i < (CO_T / 2) * IG_P; i += NT) {
            const int cl = i / IG_P, pl = i - cl * IG_P;
            const int c = co0 + cl, pos = p0 + pl;
            if (c >= Chalf || pos >= Ptot) continue;
            const int b = pos / p.Tout, t = pos - b * p.Tout;
            float a = sgate[cl][pl] + (p.bias ? __ldg(p.bias + c) : 0.f);
            float g = sgate[cl + CO_T / 2][pl] + (p.bias ? __ldg(p.bias + Chalf + c) : 0.f);
            float v;
            if (p.post == 1) {
                v = a * (1.f / (1.f + expf(-g)));                                         // GLU
            } else {
                if (p.cond) { a += p.cond[(size_t)b * p.cond_ld + c]; g += p.cond[(size_t)b * p.cond_ld + Chalf + c]; }
                v = tanhf(a) * (1.f / (1.f + expf(-g)));                                  // WN gate
            }
            const size_t o = ((size_t)b * Chalf + c) * p.Tout + t;
            if (p.resid) v += p.resid[o];
            v *= p.out_scale;
            if (p.mask) v *= p.mask[(size_t)b * p.Tout + t];
            p.y[o] = p.accumulate ? p.y[o] + v : v;
        }
    }
}

// Direct form for the ResBlock1 convolutions of the waveform branch (stride 1, kernel 3 / 7 / 11, dilation 1 / 3 / 5, "same" padding,
// 32 - 192 channels; vq2.py:723-729 -> modules.py:224-318): 90 of the ~150 convolutions of an encode and its critical path.  The
// implicit-GEMM kernels above rebuild the im2col tile element by element -- (ci, k) bookkeeping, bounds tests and a 4-byte cp.async per
// element, every element fetched K times -- so on these layers FFMA was 28 % of the instruction stream (profiles/r1n_conv_pipe_ncu_full.txt).
// Here a CTA stages, per 16 (K = 11: 8) input channels, the input WINDOW [16][32 J + (K-1) DIL] once (each sample fetched once, leaky ReLU applied as
// it lands) and the weight slab [32 channels][16 K]; a thread owns 8 output channels x J positions (t = tx + 32 j: conflict-free scalar
// reads of the window, broadcast reads of the weights) and runs the fully unrolled (k) loop: 8 + J shared-memory loads per 8 J FFMAs.  Stages are double-buffered with cp.async.  The accumulation order r = ci * K + k is the implicit GEMM's: bit-identical output.
// input channels per stage: 16, or 8 for the K = 11 layers (68 KB of staging at 16 allowed only 3 CTAs per SM)
template <int K> struct DcStage { static constexpr int CI = K >= 11 ? 8 : 16; };
template <int K, int DIL, int J>
struct DirectConv {
    static constexpr int DC_CI = DcStage<K>::CI;
    static constexpr int TP = 32 * J;                     // positions per CTA
    static constexpr int W = TP + (K - 1) * DIL;          // input window per channel
    static constexpr int WP = (W + 3) & ~3;               // row pitch
    static constexpr int RK = DC_CI * K;                  // weight columns (ci, k) per stage
    static constexpr int WPITCH = RK + 1;                 // weight slab [32 co][RK] (+1: the 8 rows a thread reads sit in different banks)
    static constexpr int XS = DC_CI * WP, WS = (32 * WPITCH + 3) & ~3;
    static constexpr size_t kSmem = 2 * (size_t)(XS + WS) * sizeof(float);
};

template <int K, int DIL, int J>
__global__ void __launch_bounds__(128) conv1d_direct_kernel(const ConvParams p) {
    using C = DirectConv<K, DIL, J>;
    extern __shared__ __align__(16) float dc_smem[];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int b = blockIdx.z, t0 = blockIdx.x * C::TP, co0 = blockIdx.y * 32;
    const float* xb = p.x + (size_t)b * p.Cin * p.Tin;
    const int in0 = t0 - p.pad;
    const int nst = (p.Cin + C::DC_CI - 1) / C::DC_CI;
    const bool lrelu = p.pre_lrelu != 0;

    auto issue = [&](int s) {
        float* xs = dc_smem + (s & 1) * (C::XS + C::WS);
        float* ws = xs + C::XS;
        const int c0 = s * C::DC_CI;
        for (int i = tid; i < C::DC_CI * C::W; i += 128) {
            const int ci = i / C::W, u = i - ci * C::W;
            const int c = c0 + ci, ti = in0 + u;
            const bool ok = c < p.Cin && ti >= 0 && ti < p.Tin;
            cp_async4(xs + ci * C::WP + u, ok ? xb + (size_t)c * p.Tin + ti : p.x, ok);
        }
        // weights: lanes run along (ci, k), which is contiguous in w[co][ci][k] -> one cache line per warp copy (with lanes along co every
        // copy touched 32 lines and the L1 tag stage, not the FMA pipe, set the pace: r1u, 188 us for a layer the FMAs need 60 us for)
        for (int i = tid; i < 32 * C::RK; i += 128) {
            const int co = i / C::RK, rk = i - co * C::RK;
            const bool ok = c0 * K + rk < p.Cin * K && co0 + co < p.Cout;
            cp_async4(ws + co * C::WPITCH + rk, ok ? p.w + ((size_t)(co0 + co) * p.Cin + c0) * K + rk : p.w, ok);
        }
    };

    float acc[8][J];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < J; ++j) acc[i][j] = 0.f;

    issue(0);
    cp_async_commit();
    for (int s = 0; s < nst; ++s) {
        cp_async_wait<0>();                                  // stage s has landed (this thread's copies)
        float* xs = dc_smem + (s & 1) * (C::XS + C::WS);
        if (lrelu) {                                         // once per sample, by the thread that copied it
            for (int i = tid; i < C::DC_CI * C::W; i += 128) {
                const int ci = i / C::W, u = i - ci * C::W;
                float* e = xs + ci * C::WP + u;
                const float v = *e;
                *e = v > 0.f ? v : 0.1f * v;
            }
        }
        __syncthreads();                                     // everybody's copies are visible; everybody is done with stage s - 1
        if (s + 1 < nst) issue(s + 1);                       // into the buffer stage s - 1 used
        cp_async_commit();
        const float* ws = xs + C::XS;
        const int cimax = min(C::DC_CI, p.Cin - s * C::DC_CI);
        for (int ci = 0; ci < cimax; ++ci) {
            const float* xr = xs + ci * C::WP + tx;
            const float* wr = ws + (ty * 8) * C::WPITCH + ci * K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float wa[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) wa[i] = wr[i * C::WPITCH + k];          // warp-uniform addresses: broadcast reads
                float xv[J];
#pragma unroll
                for (int j = 0; j < J; ++j) xv[j] = xr[32 * j + k * DIL];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < J; ++j) acc[i][j] = fmaf(wa[i], xv[j], acc[i][j]);
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + ty * 8 + i;
        if (co >= p.Cout) continue;
        const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int t = t0 + tx + 32 * j;
            if (t >= p.Tout) continue;
            float v = acc[i][j] + bv;
            const size_t o = ((size_t)b * p.Cout + co) * p.Tout + t;
            if (p.resid) v += p.resid[o];
            v *= p.out_scale;
            p.y[o] = p.accumulate ? p.y[o] + v : v;
        }
    }
}

template <int K, int DIL, int J>
static int conv1d_direct_launch(const ConvParams& p, cudaStream_t st) {
    using C = DirectConv<K, DIL, J>;
    static bool attr = false;
    if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(conv1d_direct_kernel<K, DIL, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem)); attr = true; }
    dim3 grid((p.Tout + C::TP - 1) / C::TP, (p.Cout + 31) / 32, p.B);
    conv1d_direct_kernel<K, DIL, J><<<grid, 128, C::kSmem, st>>>(p);
    TTTS_LAUNCH_CHECK("conv1d_direct");
    return TTTS_OK;
}

template <int J>
static int conv1d_direct_dispatch(const ConvParams& p, cudaStream_t st) {
#define TTTS_DC(KK, DD) if (p.K == KK && p.dil == DD) return conv1d_direct_launch<KK, DD, J>(p, st)
    TTTS_DC(3, 1); TTTS_DC(3, 3); TTTS_DC(3, 5);
    TTTS_DC(7, 1); TTTS_DC(7, 3); TTTS_DC(7, 5);
    TTTS_DC(11, 1); TTTS_DC(11, 3); TTTS_DC(11, 5);
#undef TTTS_DC
    return -1;
}

// returns -1 when the layer is not one the direct kernel covers (the caller falls through to the implicit GEMM)
static int conv1d_direct_try(const ConvParams& p, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_CONV_DIRECT"); on = (e && e[0] == '0') ? 0 : 1; }
    if (!on || p.stride != 1 || p.post != 0 || p.mask != nullptr || p.Tout != p.Tin || p.B > 65535 || p.Tout < 256) return -1;
    if (!(p.K == 3 || p.K == 7 || p.K == 11) || !(p.dil == 1 || p.dil == 3 || p.dil == 5)) return -1;
    // J = 3 (96 positions per CTA) or 4 (128): the smaller tail waste wins; the layer must still fill the machine twice over
    const int waste3 = (p.Tout + 95) / 96 * 96 - p.Tout, waste4 = (p.Tout + 127) / 128 * 128 - p.Tout;
    const bool j3 = waste3 * 128 < waste4 * 96;
    const long long ctas = (long long)((p.Tout + (j3 ? 95 : 127)) / (j3 ? 96 : 128)) * ((p.Cout + 31) / 32) * p.B;
    if (ctas < 2 * num_sms()) return -1;
    return j3 ? conv1d_direct_dispatch<3>(p, st) : conv1d_direct_dispatch<4>(p, st);
}

// weight norm: w[co, :] = g[co] * v[co, :] / ||v[co, :]||      (torch.nn.utils.weight_norm, dim=0)
__global__ void weight_norm_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w, int Cout, int n) {
    const int co = blockIdx.x;
    const float* vr = v + (size_t)co * n;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += vr[i] * vr[i];
    __shared__ float sm[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) sm[0] = g[co] / sqrtf(s);
    }
    __syncthreads();
    const float sc = sm[0];
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[(size_t)co * n + i] = vr[i] * sc;
}

// SnakeBeta through the anti-aliased Activation1d (alias_free_torch/act.py:8-28): 2x kaiser-sinc upsample (12 taps, replicate
// pad), x + sin^2(alpha x)/beta with alpha = exp(log_alpha), beta = exp(log_beta) (activations.py:62-119), 2x low-pass downsample.
// filt: the 12-tap kaiser-sinc filter (same for up and down; up output is scaled by ratio=2).
__global__ void snake_aa_kernel(const float* __restrict__ x, const float* __restrict__ log_alpha, const float* __restrict__ log_beta,
                                const float* __restrict__ filt, float* __restrict__ y, int C, int T) {
    // one block per (b, c) row; T is small (36) on the hot path
    extern __shared__ float sn[];
    const int row = blockIdx.x;
    const int c = row % C;
    const float* xr = x + (size_t)row * T;
    float* xp = sn;                    // padded input: T + 10 (pad 5 each side, replicate)
    float* up = sn + (T + 10);         // upsampled + activated: 2T, then padded replicate 5/6 for the down filter
    const float alpha = expf(log_alpha[c]), beta = expf(log_beta[c]);
    for (int i = threadIdx.x; i < T + 10; i += blockDim.x) xp[i] = xr[min(max(i - 5, 0), T - 1)];
    __syncthreads();
    // UpSample1d: conv_transpose1d(x_pad, filter*ratio, stride 2), then crop [pad_left : -pad_right], pad_left = 15, pad_right = 15
    // out_full[n] = 2 * sum_k xp[(n - k)/2] * f[k] over k with (n-k) even ; kept n in [15, 15 + 2T)
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) {
        const int n = i + 15;
        float s = 0.f;
        for (int k = 0; k < 12; ++k) {
            const int m = n - k;
            if (m >= 0 && (m & 1) == 0 && (m >> 1) < T + 10) s += xp[m >> 1] * filt[k];
        }
        s *= 2.f;
        const float sv = sinf(alpha * s);
        up[5 + i] = s + (1.f / (beta + 1e-9f)) * sv * sv;
    }
    __syncthreads();
    // DownSample1d = LowPassFilter1d(stride 2, pad replicate left 5 right 6, 12 taps)
    for (int i = threadIdx.x; i < 5; i += blockDim.x) up[i] = up[5];
    for (int i = threadIdx.x; i < 6; i += blockDim.x) up[5 + 2 * T + i] = up[5 + 2 * T - 1];
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < 12; ++k) s += up[2 * t + k] * filt[k];
        y[(size_t)row * T + t] = s;
    }
}

}  // namespace ttts

using namespace ttts;

extern "C" {

static int conv1d_run(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                      int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate,
                      const float* mask, int32_t post, const float* cond, int32_t cond_ld, int force_split, cudaStream_t st) {
    TTTS_CHECK_ARG(x && w && y, "conv1d: null pointer");
    TTTS_CHECK_ARG(B > 0 && Cin > 0 && Tin > 0 && Cout > 0 && K > 0 && stride > 0 && dil > 0 && pad >= 0, "conv1d: bad shape");
    const int Tout = (Tin + 2 * pad - dil * (K - 1) - 1) / stride + 1;
    TTTS_CHECK_ARG(Tout > 0, "conv1d: empty output");
    const bool gated = (post == 1 || post == 3);
    TTTS_CHECK_ARG(!gated || Cout % 2 == 0, "conv1d: gated post needs an even channel count");
    ConvParams p;
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.B = B; p.Cin = Cin; p.Tin = Tin; p.Cout = Cout; p.Tout = Tout; p.K = K; p.stride = stride;
    p.dil = dil; p.pad = pad; p.pre_lrelu = pre_lrelu; p.resid = resid; p.out_scale = out_scale; p.accumulate = accumulate; p.mask = mask;
    p.post = post; p.cond = cond; p.cond_ld = cond_ld;
    static int use_v1 = -1;
    if (use_v1 < 0) { const char* e = getenv("TTTS_CONV_V1"); use_v1 = (e && e[0] == '1') ? 1 : 0; }
    if (force_split) {                                             // ttts_conv1d_f32_split: the split-reduction kernel on the <32> tiling, whatever the layer
        const long long Ptot = (long long)B * Tout;
        TTTS_CHECK_ARG(force_split == 2 || force_split == 4, "conv1d split: groups must be 2 or 4");
        TTTS_CHECK_ARG(Ptot < (1ll << 31) && (long long)Cin * K < (1ll << 31), "conv1d: problem too large");
        const int ceff = gated ? Cout / 2 : Cout;
        dim3 grid((unsigned)((Ptot + IG_P - 1) / IG_P), (ceff + (gated ? 15 : 31)) / (gated ? 16 : 32));
        return conv1d_split_try(p, grid, force_split, st);
    }
    if (!use_v1) {
        const int rc_direct = conv1d_direct_try(p, st);
        if (rc_direct >= 0) return rc_direct;
        const long long Ptot = (long long)B * Tout;
        TTTS_CHECK_ARG(Ptot < (1ll << 31) && (long long)Cin * K < (1ll << 31), "conv1d: problem too large");
        const int ceff = gated ? Cout / 2 : Cout;                  // channels a CTA row tile is cut from
        // 32-channel tiles when the layer has few output channels (or few CTAs): no half-empty tiles, more CTAs in flight
        const long long ctas64 = ((Ptot + IG_P - 1) / IG_P) * ((ceff + (gated ? 31 : 63)) / (gated ? 32 : 64));
        const bool small = (gated ? ceff <= 16 : ceff <= 32) || ctas64 < 2 * num_sms();
        static int pipe = -1;
        if (pipe < 0) { const char* e = getenv("TTTS_CONV_PIPE"); pipe = (e && e[0] == '0') ? 0 : 1; }
        if (small) {
            dim3 grid((unsigned)((Ptot + IG_P - 1) / IG_P), (ceff + (gated ? 15 : 31)) / (gated ? 16 : 32));
            if (pipe) {
                const int rc_split = conv1d_split_try(p, grid, 0, st);        // latency-bound layers, only with TTTS_CONV_SPLIT=1
                if (rc_split >= 0) return rc_split;
            }
            if (pipe) conv1d_igemm_pipe_kernel<32><<<grid, 128, 0, st>>>(p);
            else conv1d_igemm_kernel<32><<<grid, 128, 0, st>>>(p);
        } else {
            dim3 grid((unsigned)((Ptot + IG_P - 1) / IG_P), (ceff + (gated ? 31 : 63)) / (gated ? 32 : 64));
            if (pipe) conv1d_igemm_pipe_kernel<64><<<grid, 256, 0, st>>>(p);
            else conv1d_igemm_kernel<64><<<grid, 256, 0, st>>>(p);
        }
        TTTS_LAUNCH_CHECK("conv1d_igemm");
        return TTTS_OK;
    }
    const int win = (CV_T - 1) * stride + (K - 1) * dil + 1;
    size_t smem = ((size_t)CV_CI * win + (size_t)CV_CI * K * CV_CO) * sizeof(float);
    if (smem < (size_t)CV_CO * CV_T * sizeof(float)) smem = (size_t)CV_CO * CV_T * sizeof(float);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "conv1d: window too large for shared memory");
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
        TTTS_CUDA(cudaFuncSetAttribute(conv1d_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    const int cg = gated ? (Cout / 2 + CV_CO / 2 - 1) / (CV_CO / 2) : (Cout + CV_CO - 1) / CV_CO;
    dim3 grid((Tout + CV_T - 1) / CV_T, cg, B);
    conv1d_f32_kernel<<<grid, CV_THREADS, smem, st>>>(p);
    TTTS_LAUNCH_CHECK("conv1d_f32");
    return TTTS_OK;
}

int ttts_conv1d_f32(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                    int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate,
                    const float* mask, int32_t post, const float* cond, int32_t cond_ld, void* stream) {
    // profiling family 4 (host_util.cu): the fp32 forward kernels -- also the input gradients that run on them (flipped weights / phases)
    const int Tout_ = (Tin + 2 * pad - dil * (K - 1) - 1) / (stride > 0 ? stride : 1) + 1;
    prof_begin(4, (cudaStream_t)stream, 2.0 * B * (double)(Tout_ > 0 ? Tout_ : 0) * Cin * Cout * K);
    const int rc = conv1d_run(x, w, bias, y, B, Cin, Tin, Cout, K, stride, dil, pad, pre_lrelu, resid, out_scale, accumulate, mask, post, cond, cond_ld, 0,
                              (cudaStream_t)stream);
    prof_end(4, (cudaStream_t)stream);
    return rc;
}
int ttts_conv1d_f32_split(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                          int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate,
                          const float* mask, int32_t post, const float* cond, int32_t cond_ld, int32_t groups, void* stream) {
    return conv1d_run(x, w, bias, y, B, Cin, Tin, Cout, K, stride, dil, pad, pre_lrelu, resid, out_scale, accumulate, mask, post, cond, cond_ld,
                      groups, (cudaStream_t)stream);
}

int ttts_weight_norm(const float* v, const float* g, float* w, int32_t Cout, int32_t n_per_out, void* stream) {
    TTTS_CHECK_ARG(v && g && w && Cout > 0 && n_per_out > 0, "weight_norm: bad args");
    weight_norm_kernel<<<Cout, 128, 0, (cudaStream_t)stream>>>(v, g, w, Cout, n_per_out);
    TTTS_LAUNCH_CHECK("weight_norm");
    return TTTS_OK;
}

int ttts_snake_aa(const float* x, const float* log_alpha, const float* log_beta, const float* filt12, float* y, int32_t B, int32_t C, int32_t T,
                  void* stream) {
    TTTS_CHECK_ARG(x && log_alpha && log_beta && filt12 && y && B > 0 && C > 0 && T > 0, "snake_aa: bad args");
    const size_t smem = ((size_t)(T + 10) + (size_t)(2 * T + 11)) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 48 * 1024, "snake_aa: T too large");
    snake_aa_kernel<<<B * C, 64, smem, (cudaStream_t)stream>>>(x, log_alpha, log_beta, filt12, y, C, T);
    TTTS_LAUNCH_CHECK("snake_aa");
    return TTTS_OK;
}
}

// ------------------------------------------------------------------------------------------------------------
// MelStyleEncoder helpers (ttts/vqvae/modules.py:600-764): tiny multi-head self-attention over T <= 64 frames with a key
// padding mask, and the masked temporal mean.  Activations are channel-major [B, C, T] (the layout the conv kernels write).
// ------------------------------------------------------------------------------------------------------------
namespace ttts {

// out[b, h*dk + j, tq] = sum_tk softmax_tk(q[:, tq] . k[:, tk] / temperature) v[j, tk]   (keys tk >= len[b] masked to -inf)
__global__ void __launch_bounds__(64) mha_small_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                       const int64_t* __restrict__ lens, float* __restrict__ out, int C, int T, int dk,
                                                       float inv_temp) {
    extern __shared__ float ms[];
    const int b = blockIdx.y, h = blockIdx.x;
    float* sq = ms;                 // [dk][T]
    float* sk = sq + dk * T;
    float* sv = sk + dk * T;
    const size_t base = ((size_t)b * C + (size_t)h * dk) * T;
    for (int i = threadIdx.x; i < dk * T; i += blockDim.x) { sq[i] = q[base + i]; sk[i] = k[base + i]; sv[i] = v[base + i]; }
    __syncthreads();
    const int len = lens ? (int)min((int64_t)T, lens[b]) : T;
    const int tq = threadIdx.x;
    if (tq >= T) return;
    float s[64];
    float mx = -INFINITY;
    for (int tk = 0; tk < T; ++tk) {
        float a = 0.f;
        for (int j = 0; j < dk; ++j) a = fmaf(sq[j * T + tq], sk[j * T + tk], a);
        a = (tk < len) ? a * inv_temp : -INFINITY;
        s[tk] = a;
        mx = fmaxf(mx, a);
    }
    float sum = 0.f;
    for (int tk = 0; tk < T; ++tk) { s[tk] = expf(s[tk] - mx); sum += s[tk]; }
    const float inv = 1.f / sum;
    for (int j = 0; j < dk; ++j) {
        float a = 0.f;
        for (int tk = 0; tk < T; ++tk) a = fmaf(s[tk], sv[j * T + tk], a);
        out[base + (size_t)j * T + tq] = a * inv;
    }
}

// y[b, c] = sum_{t < len[b]} x[b, c, t] / len[b]
__global__ void masked_mean_kernel(const float* __restrict__ x, const int64_t* __restrict__ lens, float* __restrict__ y, int C, int T) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int len = lens ? (int)min((int64_t)T, lens[b]) : T;
    float s = 0.f;
    for (int t = 0; t < len; ++t) s += x[((size_t)b * C + c) * T + t];
    y[(size_t)b * C + c] = s / (float)len;
}

// z = (m + eps * exp(logs)) * mask        (PosteriorAudioEncoder tail, vq2.py:742-744); stats = [B, 2C, T] (m | logs)
__global__ void posterior_sample_kernel(const float* __restrict__ stats, const float* __restrict__ eps, const float* __restrict__ mask,
                                        float* __restrict__ z, int C, int T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bt = i / ((size_t)C * T);
        const size_t r = i - bt * C * T;
        const int t = (int)(r % T);
        const float m = stats[bt * 2 * C * T + r];
        const float ls = stats[bt * 2 * C * T + (size_t)C * T + r];
        const float e = eps ? eps[i] : 0.f;
        z[i] = (m + e * expf(ls)) * (mask ? mask[bt * T + t] : 1.f);
    }
}

}  // namespace ttts

extern "C" {

int ttts_mha_small(const float* q, const float* k, const float* v, const int64_t* lens, float* out, int32_t B, int32_t C, int32_t T, int32_t heads,
                   float temperature, void* stream) {
    TTTS_CHECK_ARG(q && k && v && out && B > 0 && C > 0 && heads > 0 && C % heads == 0, "mha_small: bad args");
    TTTS_CHECK_ARG(T >= 1 && T <= 64, "mha_small: T must be <= 64 (got %d)", T);
    const int dk = C / heads;
    const size_t smem = (size_t)3 * dk * T * sizeof(float);
    TTTS_CHECK_ARG(smem <= 48 * 1024, "mha_small: head too large");
    ttts::mha_small_kernel<<<dim3(heads, B), 64, smem, (cudaStream_t)stream>>>(q, k, v, lens, out, C, T, dk, 1.0f / temperature);
    TTTS_LAUNCH_CHECK("mha_small");
    return TTTS_OK;
}

int ttts_masked_mean(const float* x, const int64_t* lens, float* y, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(x && y && B > 0 && C > 0 && T > 0, "masked_mean: bad args");
    ttts::masked_mean_kernel<<<dim3((C + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(x, lens, y, C, T);
    TTTS_LAUNCH_CHECK("masked_mean");
    return TTTS_OK;
}

int ttts_posterior_sample(const float* stats, const float* eps, const float* mask, float* z, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(stats && z && B > 0 && C > 0 && T > 0, "posterior_sample: bad args");
    const size_t n = (size_t)B * C * T;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ttts::posterior_sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(stats, eps, mask, z, C, T, n);
    TTTS_LAUNCH_CHECK("posterior_sample");
    return TTTS_OK;
}
}
