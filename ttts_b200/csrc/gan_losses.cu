// NOT YET RUN ON HARDWARE (validated on the CPU emulation of this source against tests/ref_kernels.py).  Next scope row:
// the reductions behind the adversarial losses of the VQ-VAE-GAN step (ttts/vqvae/losses.py:7-44):
//   least-squares GAN terms   mean((c - x)^2)      discriminator_loss: c = 1 on real logits, c = 0 on generated ; generator_loss: c = 1
//   feature matching          mean(|a - b|)        feature_loss, a = the (detached) real feature map, b = the generated one
// Two-stage, fixed grid: RED_BLOCKS per-block partials summed by one block in index order -> deterministic.  Backward is element-wise.
#include <stdlib.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#endif

namespace ttts {

constexpr int RED_BLOCKS = 256;

// mode 0: (c - x)^2 ; mode 1: |a - x|  (a = other)
__global__ void __launch_bounds__(256) loss_partial_kernel(const float* __restrict__ x, const float* __restrict__ other, float c, size_t n, int mode,
                                                           float* __restrict__ partial) {
    __shared__ float red[8];
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)RED_BLOCKS * 256) {
        const float v = x[i];
        if (mode == 0) { const float d = c - v; s = fmaf(d, d, s); }
        else s += fabsf(other[i] - v);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) loss_final_kernel(const float* __restrict__ partial, float inv_n, float* __restrict__ out) {
    __shared__ float red[8];
    float s = threadIdx.x < RED_BLOCKS ? partial[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[0] = t * inv_n;
    }
}
// dx = dL * d/dx : mode 0: 2 (x - c) / n ; mode 1: sign(x - a) / n
__global__ void loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ other, float c, const float* __restrict__ dL, size_t n, int mode,
                                float inv_n, float* __restrict__ dx) {
    const float g = dL[0] * inv_n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        if (mode == 0) dx[i] = 2.f * (v - c) * g;
        else { const float d = v - other[i]; dx[i] = d > 0.f ? g : (d < 0.f ? -g : 0.f); }
    }
}

// KL term between the flowed posterior and the prior (losses.py:47-61):
//   kl = sum((logs_p - logs_q - 0.5 + 0.5 (z_p - m_p)^2 exp(-2 logs_p)) mask) / sum(mask) ; tensors [B, C, T], mask [B, T]
__global__ void __launch_bounds__(256) kl_partial_kernel(const float* __restrict__ z_p, const float* __restrict__ logs_q, const float* __restrict__ m_p,
                                                         const float* __restrict__ logs_p, const float* __restrict__ mask, int C, int T, size_t n,
                                                         float* __restrict__ partial) {
    __shared__ float red[2][8];
    float s = 0.f, sm = 0.f;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)RED_BLOCKS * 256) {
        const size_t b = i / ((size_t)C * T);
        const int t = (int)(i % T);
        const float mk = mask[b * T + t];
        const float d = z_p[i] - m_p[i], lp = logs_p[i];
        s += (lp - logs_q[i] - 0.5f + 0.5f * d * d * expf(-2.f * lp)) * mk;
        if ((i / T) % C == 0) sm += mk;                              // the reference sums the [B, 1, T] mask: once per (b, t)
    }
    s = warp_sum(s); sm = warp_sum(sm);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = sm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
        partial[blockIdx.x] = a;
        partial[RED_BLOCKS + blockIdx.x] = c;
    }
}
// out[0] = kl ; out[1] = sum(mask) (kept for the backward)
__global__ void __launch_bounds__(256) kl_final_kernel(const float* __restrict__ partial, float* __restrict__ out) {
    __shared__ float red[2][8];
    float s = partial[threadIdx.x], sm = partial[RED_BLOCKS + threadIdx.x];
    s = warp_sum(s); sm = warp_sum(sm);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = sm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
        out[0] = a / c;
        out[1] = c;
    }
}
__global__ void kl_bwd_kernel(const float* __restrict__ z_p, const float* __restrict__ m_p, const float* __restrict__ logs_p,
                              const float* __restrict__ mask, const float* __restrict__ dL, const float* __restrict__ fwd_out, int C, int T, size_t n,
                              float* __restrict__ dz_p, float* __restrict__ dlogs_q, float* __restrict__ dm_p, float* __restrict__ dlogs_p) {
    const float g = dL[0] / fwd_out[1];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)C * T);
        const int t = (int)(i % T);
        const float gm = g * mask[b * T + t];
        const float d = z_p[i] - m_p[i], e = expf(-2.f * logs_p[i]);
        if (dz_p) dz_p[i] = gm * d * e;
        if (dlogs_q) dlogs_q[i] = -gm;
        if (dm_p) dm_p[i] = -gm * d * e;
        if (dlogs_p) dlogs_p[i] = gm * (1.f - d * d * e);
    }
}

static int loss_fwd(const float* x, const float* other, float c, int64_t n, int mode, float* scratch, float* out, cudaStream_t st) {
    TTTS_CHECK_ARG(x && scratch && out && n > 0 && (mode == 0 || other), "loss: bad args");
    TTTS_CUDA(launch_plain(loss_partial_kernel, dim3(RED_BLOCKS), dim3(256), 0, st, x, other, c, (size_t)n, mode, scratch));
    TTTS_LAUNCH_CHECK("loss_partial");
    TTTS_CUDA(launch_plain(loss_final_kernel, dim3(1), dim3(256), 0, st, (const float*)scratch, 1.0f / (float)n, out));
    TTTS_LAUNCH_CHECK("loss_final");
    return TTTS_OK;
}
static int loss_bwd(const float* x, const float* other, float c, const float* dL, int64_t n, int mode, float* dx, cudaStream_t st) {
    TTTS_CHECK_ARG(x && dL && dx && n > 0 && (mode == 0 || other), "loss backward: bad args");
    size_t b = ((size_t)n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (b > cap) b = cap;
    TTTS_CUDA(launch_plain(loss_bwd_kernel, dim3((unsigned)b), dim3(256), 0, st, x, other, c, dL, (size_t)n, mode, 1.0f / (float)n, dx));
    TTTS_LAUNCH_CHECK("loss_bwd");
    return TTTS_OK;
}

}  // namespace ttts

/* out[0] = mean((c - x)^2) ; scratch: 256 floats */
extern "C" int ttts_lsgan_loss(const float* x, float c, int64_t n, float* scratch, float* out, void* stream) {
    return ttts::loss_fwd(x, nullptr, c, n, 0, scratch, out, (cudaStream_t)stream);
}
/* dx = dL[0] * 2 (x - c) / n */
extern "C" int ttts_lsgan_loss_bwd(const float* x, float c, const float* dL, int64_t n, float* dx, void* stream) {
    return ttts::loss_bwd(x, nullptr, c, dL, n, 0, dx, (cudaStream_t)stream);
}
/* out[0] = mean(|a - b|) ; scratch: 256 floats */
extern "C" int ttts_l1_mean(const float* a, const float* b, int64_t n, float* scratch, float* out, void* stream) {
    return ttts::loss_fwd(b, a, 0.f, n, 1, scratch, out, (cudaStream_t)stream);
}
/* db = dL[0] * sign(b - a) / n   (a is the detached side) */
extern "C" int ttts_l1_mean_bwd(const float* a, const float* b, const float* dL, int64_t n, float* db, void* stream) {
    return ttts::loss_bwd(b, a, 0.f, dL, n, 1, db, (cudaStream_t)stream);
}
/* out[0] = kl_loss(z_p, logs_q, m_p, logs_p, mask) (losses.py:47-61), out[1] = sum(mask) ; tensors [B,C,T], mask [B,T] ; scratch: 512 floats */
extern "C" int ttts_kl_loss(const float* z_p, const float* logs_q, const float* m_p, const float* logs_p, const float* mask, int32_t B, int32_t C,
                            int32_t T, float* scratch, float* out2, void* stream) {
    TTTS_CHECK_ARG(z_p && logs_q && m_p && logs_p && mask && scratch && out2 && B > 0 && C > 0 && T > 0, "kl_loss: bad args");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(ttts::launch_plain(ttts::kl_partial_kernel, dim3(ttts::RED_BLOCKS), dim3(256), 0, (cudaStream_t)stream, z_p, logs_q, m_p, logs_p, mask, C, T,
                                 n, scratch));
    TTTS_LAUNCH_CHECK("kl_partial");
    TTTS_CUDA(ttts::launch_plain(ttts::kl_final_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, (const float*)scratch, out2));
    TTTS_LAUNCH_CHECK("kl_final");
    return TTTS_OK;
}
/* gradients of the four inputs (any of the outputs may be NULL); fwd_out2 = what ttts_kl_loss wrote */
extern "C" int ttts_kl_loss_bwd(const float* z_p, const float* m_p, const float* logs_p, const float* mask, const float* dL, const float* fwd_out2,
                                int32_t B, int32_t C, int32_t T, float* dz_p, float* dlogs_q, float* dm_p, float* dlogs_p, void* stream) {
    TTTS_CHECK_ARG(z_p && m_p && logs_p && mask && dL && fwd_out2 && B > 0 && C > 0 && T > 0, "kl_loss backward: bad args");
    const size_t n = (size_t)B * C * T;
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)ttts::num_sms() * 8;
    if (b > cap) b = cap;
    TTTS_CUDA(ttts::launch_plain(ttts::kl_bwd_kernel, dim3((unsigned)b), dim3(256), 0, (cudaStream_t)stream, z_p, m_p, logs_p, mask, dL, fwd_out2, C, T, n,
                                 dz_p, dlogs_q, dm_p, dlogs_p));
    TTTS_LAUNCH_CHECK("kl_bwd");
    return TTTS_OK;
}
