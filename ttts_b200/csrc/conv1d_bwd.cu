// NOT YET RUN ON HARDWARE (written after the round's GPU budget was spent; validated on the CPU emulation of this source, tests/emu,
// against torch.autograd).  First kernels of the next scope row (the VQ-VAE-GAN train step): the backward of the fp32
// Conv1d that ttts_conv1d_f32 computes -- input gradient, weight gradient, bias gradient -- for any stride / dilation / padding, i.e. the
// autograd of nn.Conv1d in PosteriorAudioEncoder / WN / ResBlock1 / the downsampling stack (ttts/vqvae/vq2.py:667-745, modules.py:136-318).
// Exact fp32 on CUDA cores like the forward; correctness first: double-buffered register staging, no cp.async pipeline yet.
//
//   dgrad:  dx[b, ci, ti] (+)= lrelu'(x[b, ci, ti]) * sum_{co, k} w[co, ci, k] * dy[b, co, (ti + pad - k dil) / stride]   (when divisible, in range)
//           implicit GEMM: rows = 32 input channels, columns = 64 input positions (batch x time flattened), reduction r = co * K + k
//   wgrad:  dw[co, ci, k] += sum_{b, to} dy[b, co, to] * lrelu(x)[b, ci, to stride + k dil - pad]
//           implicit GEMM: rows = 32 output channels, columns = 64 of the (ci, k) pairs (= dw's memory order), reduction over the B * Tout
//           positions, cut into slices across grid.z; slices combine with fp32 atomic adds (gradients ACCUMULATE, as in the GPT engine)
//   bgrad:  db[co] += sum_{b, to} dy[b, co, to]                                                    one CTA per channel, fixed order
#include <stdlib.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#define TTTS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif
#include "conv_params.h"

namespace ttts {

struct ConvBwdParams {
    const float* dy;          // [B, Cout, Tout]
    const float* w;           // [Cout, Cin, K]
    const float* x;           // [B, Cin, Tin] forward input (wgrad: always; dgrad: only with pre_lrelu)
    float* dx;                // [B, Cin, Tin]
    float* dw;                // [Cout, Cin, K]
    int B, Cin, Tin, Cout, Tout, K, stride, dil, pad;
    int pre_lrelu;            // the forward applied leaky_relu(0.1) to x before the convolution
    int accumulate;           // dgrad: dx += instead of dx =
    int slice;                // wgrad: positions per grid.z slice (a multiple of IG_R)
};

constexpr int BW_T = 32, BW_NT = BW_T * 4, BW_LDA = BW_T + 4, BW_LDB = IG_P + 4, BW_NB = IG_R * IG_P / BW_NT;

// ------------------------------------------------------------------------------------------------------------
// input gradient
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_NT) conv1d_dgrad_kernel(const ConvBwdParams p) {
    __shared__ __align__(16) float sA[2][IG_R][BW_LDA];
    __shared__ __align__(16) float sB[2][IG_R][BW_LDB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int p0 = blockIdx.x * IG_P, ci0 = blockIdx.y * BW_T;
    const int R = p.Cout * p.K, Ptot = p.B * p.Tin;
    const int nchunks = (R + IG_R - 1) / IG_R;
    // B loader: column cb = one input position, rows rb0 + 2 i
    const int cb = tid & 63, rb0 = tid >> 6;
    const int posb = p0 + cb;
    const bool pos_ok = posb < Ptot;
    const int bb = pos_ok ? posb / p.Tin : 0;
    const int tin = pos_ok ? posb - bb * p.Tin : 0;
    const float* dyb = p.dy + (size_t)bb * p.Cout * p.Tout;
    // A loader: column ca = input channel, rows ra4 .. ra4 + 3
    const int ca = tid >> 2, ra4 = (tid & 3) * 4;
    const bool cia_ok = ci0 + ca < p.Cin;

    float ra[4], rb[BW_NB];
    auto gload = [&](int chunk) {
        const int r0 = chunk * IG_R;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = r0 + ra4 + i;
            const int co = rr / p.K, k = rr - co * p.K;
            ra[i] = (cia_ok && rr < R) ? __ldg(p.w + ((size_t)co * p.Cin + ci0 + ca) * p.K + k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BW_NB; ++i) {
            const int rr = r0 + rb0 + (BW_NT / 64) * i;
            const int co = rr / p.K, k = rr - co * p.K;
            const int num = tin + p.pad - k * p.dil;
            float v = 0.f;
            if (pos_ok && rr < R && num >= 0) {
                const int to = num / p.stride;
                if (to * p.stride == num && to < p.Tout) v = __ldg(dyb + (size_t)co * p.Tout + to);
            }
            rb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sA[buf][ra4 + i][ca] = ra[i];
#pragma unroll
        for (int i = 0; i < BW_NB; ++i) sB[buf][rb0 + (BW_NT / 64) * i][cb] = rb[i];
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int pos = p0 + tx * 4 + j;
        if (pos >= Ptot) continue;
        const int b = pos / p.Tin, t = pos - b * p.Tin;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ci = ci0 + ty * 4 + i;
            if (ci >= p.Cin) continue;
            const size_t o = ((size_t)b * p.Cin + ci) * p.Tin + t;
            float v = acc[i][j];
            if (p.pre_lrelu) v *= p.x[o] > 0.f ? 1.f : 0.1f;
            p.dx[o] = p.accumulate ? p.dx[o] + v : v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_NT) conv1d_wgrad_kernel(const ConvBwdParams p) {
    __shared__ __align__(16) float sA[2][IG_R][BW_LDA];
    __shared__ __align__(16) float sB[2][IG_R][BW_LDB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * IG_P, co0 = blockIdx.y * BW_T;
    const int N = p.Cin * p.K, Ptot = p.B * p.Tout;
    const int q_begin = blockIdx.z * p.slice, q_end = min(Ptot, q_begin + p.slice);
    const int nchunks = (max(0, q_end - q_begin) + IG_R - 1) / IG_R;
    if (nchunks == 0) return;
    // B loader: column cb = one (ci, k) pair, rows (positions) rb0 + 2 i
    const int cb = tid & 63, rb0 = tid >> 6;
    const int nb = n0 + cb;
    const bool n_ok = nb < N;
    const int cib = n_ok ? nb / p.K : 0;
    const int kb = n_ok ? nb - cib * p.K : 0;
    const int toff = kb * p.dil - p.pad;
    // A loader: column ca = output channel, rows ra4 .. ra4 + 3
    const int ca = tid >> 2, ra4 = (tid & 3) * 4;
    const bool coa_ok = co0 + ca < p.Cout;

    float ra[4], rb[BW_NB];
    auto gload = [&](int chunk) {
        const int q0 = q_begin + chunk * IG_R;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + ra4 + i;
            float v = 0.f;
            if (coa_ok && q < q_end) { const int b = q / p.Tout, to = q - b * p.Tout; v = __ldg(p.dy + ((size_t)b * p.Cout + co0 + ca) * p.Tout + to); }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < BW_NB; ++i) {
            const int q = q0 + rb0 + (BW_NT / 64) * i;
            float v = 0.f;
            if (n_ok && q < q_end) {
                const int b = q / p.Tout, to = q - b * p.Tout;
                const int ti = to * p.stride + toff;
                if (ti >= 0 && ti < p.Tin) {
                    v = __ldg(p.x + ((size_t)b * p.Cin + cib) * p.Tin + ti);
                    if (p.pre_lrelu) v = v > 0.f ? v : 0.1f * v;
                }
            }
            rb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sA[buf][ra4 + i][ca] = ra[i];
#pragma unroll
        for (int i = 0; i < BW_NB; ++i) sB[buf][rb0 + (BW_NT / 64) * i][cb] = rb[i];
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) atomicAdd(p.dw + (size_t)co * N + n, acc[i][j]);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// weight gradient, pipelined (default; TTTS_WGRAD_V1=1 = the kernel above).  r2m launch list of the diffusion step: conv1d_wgrad_kernel 37 % of the
// step at ~12 TFLOP/s while the forward implicit GEMM runs the same FLOPs at ~40.  Same GEMM (rows = output channels, columns = (ci, k) pairs,
// reduction over the B * Tout positions, slices across grid.z combined by fp32 atomic adds), but: 64 x 64 tile and 256 threads (4 x 4 outputs per
// thread, strided by 16 so that every float4 shared-memory read of a quarter-warp is conflict-free), chunks of 32 positions staged AS THEY LIE in
// global memory (position-contiguous rows: [channel][32 positions], pitch 36) by 4-byte cp.async with zero fill through a 3-stage ring, the
// reduction index inside a chunk vectorised (8 LDS.128 per 64 FMA), the leaky ReLU applied once per element by the thread that copied it.
// ------------------------------------------------------------------------------------------------------------
constexpr int W2_T = 64, W2_R = 32, W2_LD = W2_R + 4, W2_STAGES = 3, W2_NT = 256;
constexpr int W2_SMEM = W2_STAGES * 2 * W2_T * W2_LD * (int)sizeof(float);

__global__ void __launch_bounds__(W2_NT) conv1d_wgrad2_kernel(const ConvBwdParams p) {
    TTTS_DYN_SMEM(float, w2_smem);
    float* sA = w2_smem;                                   // [stage][co][W2_LD]
    float* sB = w2_smem + W2_STAGES * W2_T * W2_LD;        // [stage][n][W2_LD]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * W2_T, co0 = blockIdx.y * W2_T;
    const int N = p.Cin * p.K, Ptot = p.B * p.Tout;
    const int q_begin = blockIdx.z * p.slice, q_end = min(Ptot, q_begin + p.slice);
    const int nchunks = (max(0, q_end - q_begin) + W2_R - 1) / W2_R;
    if (nchunks == 0) return;
    // loader: this thread copies position rr of rows r0, r0 + 8, ... of both operand tiles
    const int rr = tid & 31, r0 = tid >> 5;
    int xoff[8];                                           // ci * Tin + k * dil - pad of the eight (ci, k) columns, or a negative flag
    int toffv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + r0 + 8 * i;
        if (n < N) { const int ci = n / p.K, k = n - ci * p.K; xoff[i] = ci * p.Tin; toffv[i] = k * p.dil - p.pad; }
        else { xoff[i] = -1; toffv[i] = 0; }
    }
    auto issue = [&](int chunk) {
        const int st = chunk % W2_STAGES;
        const int q = q_begin + chunk * W2_R + rr;
        const bool q_ok = q < q_end;
        const int b = q_ok ? q / p.Tout : 0, to = q_ok ? q - b * p.Tout : 0;
        const float* dyb = p.dy + (size_t)b * p.Cout * p.Tout + to;
        const float* xb = p.x + (size_t)b * p.Cin * p.Tin;
        float* a = sA + (st * W2_T + r0) * W2_LD + rr;
        float* bs = sB + (st * W2_T + r0) * W2_LD + rr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int co = co0 + r0 + 8 * i;
            const bool ok = q_ok && co < p.Cout;
            cp_async4(a + 8 * i * W2_LD, ok ? dyb + (size_t)co * p.Tout : p.dy, ok);
            const int ti = to * p.stride + toffv[i];
            const bool okb = q_ok && xoff[i] >= 0 && ti >= 0 && ti < p.Tin;
            cp_async4(bs + 8 * i * W2_LD, okb ? xb + xoff[i] + ti : p.x, okb);
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < W2_STAGES - 1; ++c) { if (c < nchunks) issue(c); cp_async_commit(); }
    for (int c = 0; c < nchunks; ++c) {
        const int st = c % W2_STAGES;
        cp_async_wait<W2_STAGES - 2>();
        if (p.pre_lrelu) {                                  // this thread's own copies of chunk c have landed: activate them in place
            float* bs = sB + (st * W2_T + r0) * W2_LD + rr;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float v = bs[8 * i * W2_LD]; bs[8 * i * W2_LD] = v > 0.f ? v : 0.1f * v; }
        }
        __syncthreads();
        if (c + W2_STAGES - 1 < nchunks) issue(c + W2_STAGES - 1);
        cp_async_commit();
        const float* a = sA + st * W2_T * W2_LD;
        const float* bs = sB + st * W2_T * W2_LD;
#pragma unroll
        for (int r = 0; r < W2_R; r += 4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a + (ty + 16 * i) * W2_LD + r);
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(bs + (tx + 16 * j) * W2_LD + r);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] = fmaf(av[i].x, bv[j].x, fmaf(av[i].y, bv[j].y, fmaf(av[i].z, bv[j].z, fmaf(av[i].w, bv[j].w, acc[i][j]))));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty + 16 * i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) atomicAdd(p.dw + (size_t)co * N + n, acc[i][j]);
        }
    }
}

// db[co] += sum over (b, to) of dy ; grid (Cout, slices): with one CTA per channel (the first version) a 32-channel layer over 1.3 M positions
// ran on 32 SMs (conv1d_bgrad_kernel: 7 % of the VQ-VAE-GAN step, r2o launch list).  One slice = fixed summation order; several slices combine
// by fp32 atomic adds like the weight gradient.
__global__ void __launch_bounds__(256) conv1d_bgrad_kernel(const float* __restrict__ dy, float* __restrict__ db, int B, int Cout, int Tout) {
    __shared__ float red[8];
    const int co = blockIdx.x;
    const int P = B * Tout, per = (P + (int)gridDim.y - 1) / (int)gridDim.y;
    const int q0 = blockIdx.y * per, q1 = min(P, q0 + per);
    float s = 0.f;
    for (int i = q0 + threadIdx.x; i < q1; i += 256) {
        const int b = i / Tout, t = i - b * Tout;
        s += dy[((size_t)b * Cout + co) * Tout + t];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (gridDim.y == 1) db[co] += t; else atomicAdd(db + co, t);
    }
}
// The same sum with 16-byte loads, four of them in flight per thread, for rows that allow it (Tout a multiple of 4 keeps every (b, co) row
// 16-byte aligned): an item = 1024 consecutive positions of one (b, co) row, slice y takes items y, y + gridDim.y, ...  r2ah launch list: the
// scalar kernel above read a 84 MB dy in 92 us (0.9 TB/s: one 4-byte load per thread and iteration, an integer division per element,
// 384 CTAs) -- as long as one of the GEMMs of the layer it belongs to.
__global__ void __launch_bounds__(256) conv1d_bgrad4_kernel(const float* __restrict__ dy, float* __restrict__ db, int B, int Cout, int Tout) {
    __shared__ float red[8];
    const int co = blockIdx.x;
    const int chunks = (Tout + 1023) / 1024;
    const int items = B * chunks;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int it0 = blockIdx.y; it0 < items; it0 += 4 * (int)gridDim.y) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * (int)gridDim.y;
            const int b = it / chunks, t = (it - b * chunks) * 1024 + 4 * (int)threadIdx.x;
            v[u] = (it < items && t < Tout) ? *reinterpret_cast<const float4*>(dy + ((size_t)b * Cout + co) * Tout + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { s0 += v[u].x; s1 += v[u].y; s2 += v[u].z; s3 += v[u].w; }
    }
    float s = warp_sum((s0 + s1) + (s2 + s3));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (gridDim.y == 1) db[co] += t; else atomicAdd(db + co, t);
    }
}
static inline dim3 bgrad_grid(int B, int Cout, int Tout) {
    const long long P = (long long)B * Tout;
    long long S = (2ll * num_sms() + Cout - 1) / Cout;
    const long long s_max = (P + 8191) / 8192;
    if (S > s_max) S = s_max;
    if (S < 1) S = 1;
    return dim3(Cout, (unsigned)S);
}
// db += column sums of dy [B, Cout, Tout]
static int bgrad_launch(const float* dy, float* db, int B, int Cout, int Tout, cudaStream_t st) {
    if ((Tout & 3) == 0 && (((uintptr_t)dy) & 15) == 0) {
        const long long items = (long long)B * ((Tout + 1023) / 1024);
        long long S = (8ll * num_sms() + Cout - 1) / Cout;               // ~8 CTAs per SM in all
        const long long s_max = (items + 3) / 4;                          // at least one full round of four loads per slice
        if (S > s_max) S = s_max;
        if (S < 1) S = 1;
        if (S > 65535) S = 65535;
        return (int)launch_plain(conv1d_bgrad4_kernel, dim3(Cout, (unsigned)S), dim3(256), 0, st, dy, db, B, Cout, Tout);
    }
    return (int)launch_plain(conv1d_bgrad_kernel, bgrad_grid(B, Cout, Tout), dim3(256), 0, st, dy, db, B, Cout, Tout);
}

static int bwd_params(ConvBwdParams& p, int B, int Cin, int Tin, int Cout, int K, int stride, int dil, int pad) {
    TTTS_CHECK_ARG(B > 0 && Cin > 0 && Tin > 0 && Cout > 0 && K > 0 && stride > 0 && dil > 0 && pad >= 0, "conv1d backward: bad shape");
    const int Tout = (Tin + 2 * pad - dil * (K - 1) - 1) / stride + 1;
    TTTS_CHECK_ARG(Tout > 0, "conv1d backward: empty output");
    TTTS_CHECK_ARG((long long)B * Tin < (1ll << 31) && (long long)Cin * K < (1ll << 31) && (long long)Cout * K < (1ll << 31),
                   "conv1d backward: problem too large");
    p.B = B; p.Cin = Cin; p.Tin = Tin; p.Cout = Cout; p.Tout = Tout; p.K = K; p.stride = stride; p.dil = dil; p.pad = pad;
    return TTTS_OK;
}

int conv1d_bwd_input(const float* dy, const float* w, const float* x, float* dx, int B, int Cin, int Tin, int Cout, int K, int stride, int dil,
                     int pad, int pre_lrelu, int accumulate, cudaStream_t st) {
    ConvBwdParams p = {};
    TTTS_RUN(bwd_params(p, B, Cin, Tin, Cout, K, stride, dil, pad));
    TTTS_CHECK_ARG(dy && w && dx && (!pre_lrelu || x), "conv1d dgrad: null pointer");
    p.dy = dy; p.w = w; p.x = x; p.dx = dx; p.pre_lrelu = pre_lrelu; p.accumulate = accumulate;
    const dim3 grid((unsigned)(((long long)B * Tin + IG_P - 1) / IG_P), (Cin + BW_T - 1) / BW_T);
    prof_begin(2, st, 2.0 * B * p.Tout * (double)Cin * Cout * K);
    TTTS_CUDA(launch_plain(conv1d_dgrad_kernel, grid, dim3(BW_NT), 0, st, p));
    prof_end(2, st);
    TTTS_LAUNCH_CHECK("conv1d_dgrad");
    return TTTS_OK;
}

int conv1d_bwd_weight(const float* dy, const float* x, float* dw, float* db, int B, int Cin, int Tin, int Cout, int K, int stride, int dil, int pad,
                      int pre_lrelu, cudaStream_t st) {
    ConvBwdParams p = {};
    TTTS_RUN(bwd_params(p, B, Cin, Tin, Cout, K, stride, dil, pad));
    TTTS_CHECK_ARG(dy && x && dw, "conv1d wgrad: null pointer");
    p.dy = dy; p.x = x; p.dw = dw; p.pre_lrelu = pre_lrelu;
    static int use_v1 = -1;
    if (use_v1 < 0) { const char* e = getenv("TTTS_WGRAD_V1"); use_v1 = (e && e[0] == '1') ? 1 : 0; }
    const long long Ptot = (long long)B * p.Tout;
    if (!use_v1) {
        const int gx = (Cin * K + W2_T - 1) / W2_T, gy = (Cout + W2_T - 1) / W2_T;
        // slices of the position axis: about two CTAs per SM in total, at least 8 chunks each
        long long S = (2ll * num_sms() + (long long)gx * gy - 1) / ((long long)gx * gy);
        const long long s_max = (Ptot + 8 * W2_R - 1) / (8 * W2_R);
        if (S > s_max) S = s_max;
        if (S < 1) S = 1;
        if (S > 65535) S = 65535;
        long long slice = (Ptot + S - 1) / S;
        slice = (slice + W2_R - 1) / W2_R * W2_R;
        p.slice = (int)slice;
        const dim3 grid(gx, gy, (unsigned)((Ptot + slice - 1) / slice));
#ifndef TTTS_HOST_EMU
        static bool attr = false;
        if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(conv1d_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W2_SMEM)); attr = true; }
#endif
        prof_begin(3, st, 2.0 * B * p.Tout * (double)Cin * Cout * K);
        TTTS_CUDA(launch_plain(conv1d_wgrad2_kernel, grid, dim3(W2_NT), (size_t)W2_SMEM, st, p));
        prof_end(3, st);
    } else {
    const int gx = (Cin * K + IG_P - 1) / IG_P, gy = (Cout + BW_T - 1) / BW_T;
    // slices of the position axis: about two CTAs per SM in total, at least 4 chunks each
    long long S = (2ll * num_sms() + (long long)gx * gy - 1) / ((long long)gx * gy);
    const long long s_max = (Ptot + 4 * IG_R - 1) / (4 * IG_R);
    if (S > s_max) S = s_max;
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
    long long slice = (Ptot + S - 1) / S;
    slice = (slice + IG_R - 1) / IG_R * IG_R;
    p.slice = (int)slice;
    const dim3 grid(gx, gy, (unsigned)((Ptot + slice - 1) / slice));
    prof_begin(3, st, 2.0 * B * p.Tout * (double)Cin * Cout * K);
    TTTS_CUDA(launch_plain(conv1d_wgrad_kernel, grid, dim3(BW_NT), 0, st, p));
    prof_end(3, st);
    }
    TTTS_LAUNCH_CHECK("conv1d_wgrad");
    if (db) {
        TTTS_CUDA((cudaError_t)bgrad_launch(dy, db, B, Cout, p.Tout, st));
        TTTS_LAUNCH_CHECK("conv1d_bgrad");
    }
    return TTTS_OK;
}

}  // namespace ttts

extern "C" {
int ttts_conv1d_bwd_input(const float* dy, const float* w, const float* x, float* dx, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                          int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, int32_t accumulate, void* stream) {
    return ttts::conv1d_bwd_input(dy, w, x, dx, B, Cin, Tin, Cout, K, stride, dil, pad, pre_lrelu, accumulate, (cudaStream_t)stream);
}
int ttts_bias_grad(const float* dy, float* db, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(dy && db && B > 0 && C > 0 && T > 0, "bias_grad: bad args");
    TTTS_CUDA((cudaError_t)ttts::bgrad_launch(dy, db, B, C, T, (cudaStream_t)stream));
    TTTS_LAUNCH_CHECK("bias_grad");
    return TTTS_OK;
}
int ttts_conv1d_bwd_weight(const float* dy, const float* x, float* dw, float* db, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                           int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, void* stream) {
    return ttts::conv1d_bwd_weight(dy, x, dw, db, B, Cin, Tin, Cout, K, stride, dil, pad, pre_lrelu, (cudaStream_t)stream);
}
}
