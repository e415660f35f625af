// NOT YET RUN ON HARDWARE (validated on the CPU emulation of this source against torch).  Next scope row: grouped Conv1d,
// forward and backward, for the scale discriminator of the VQ-VAE-GAN step (DiscriminatorS, ttts/vqvae/vq2.py:498-507: kernel 41, stride 4,
// groups 4 / 16 / 64 / 256, i.e. FOUR input channels per group in every layer).  With so few channels per group there is no GEMM to speak of
// (164 multiply-adds per output): one thread per output element, coalesced along time; the weight gradient is a per-(co, ci) reduction over
// batch x time with the K tap sums kept per thread and combined in a fixed order.  Correctness first.
#include <stdlib.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#endif

namespace ttts {

constexpr int GC_KMAX = 64;

struct GConvParams {
    const float *x, *w, *bias, *dy;
    float *y, *dx, *dw;
    int B, Cin, Tin, Cout, Tout, K, stride, pad, groups;
};

// y[b, co, to] = bias[co] + sum_{c < Cin/G, k} w[co, c, k] x[b, g Cin/G + c, to stride + k - pad]
__global__ void gconv_fwd_kernel(const GConvParams p) {
    const int cg = p.Cin / p.groups, og = p.Cout / p.groups;
    const size_t n = (size_t)p.B * p.Cout * p.Tout;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int to = (int)(i % p.Tout);
        const int co = (int)((i / p.Tout) % p.Cout);
        const int b = (int)(i / ((size_t)p.Tout * p.Cout));
        const int g = co / og;
        float s = p.bias ? p.bias[co] : 0.f;
        for (int c = 0; c < cg; ++c) {
            const float* xr = p.x + ((size_t)b * p.Cin + g * cg + c) * p.Tin;
            const float* wr = p.w + ((size_t)co * cg + c) * p.K;
            for (int k = 0; k < p.K; ++k) {
                const int ti = to * p.stride + k - p.pad;
                if (ti >= 0 && ti < p.Tin) s = fmaf(wr[k], xr[ti], s);
            }
        }
        p.y[i] = s;
    }
}
// dx[b, ci, ti] = sum_{co in the group of ci, k : (ti + pad - k) % stride == 0} w[co, ci_g, k] dy[b, co, (ti + pad - k) / stride]
__global__ void gconv_dgrad_kernel(const GConvParams p) {
    const int cg = p.Cin / p.groups, og = p.Cout / p.groups;
    const size_t n = (size_t)p.B * p.Cin * p.Tin;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ti = (int)(i % p.Tin);
        const int ci = (int)((i / p.Tin) % p.Cin);
        const int b = (int)(i / ((size_t)p.Tin * p.Cin));
        const int g = ci / cg, c = ci - g * cg;
        float s = 0.f;
        for (int o = 0; o < og; ++o) {
            const int co = g * og + o;
            const float* dr = p.dy + ((size_t)b * p.Cout + co) * p.Tout;
            const float* wr = p.w + ((size_t)co * cg + c) * p.K;
            // only the taps k = (ti + pad) mod stride, + stride, ... meet this position: to = (ti + pad - k) / stride falls by one per step
            // (the first version tested all K taps with a division each: gconv_dgrad_kernel was 14 % of the VQ-VAE-GAN step, r2o launch list)
            int k = (ti + p.pad) % p.stride;
            int to = (ti + p.pad - k) / p.stride;
            for (; k < p.K && to >= 0; k += p.stride, --to)
                if (to < p.Tout) s = fmaf(wr[k], dr[to], s);
        }
        p.dx[i] = s;
    }
}
// dw[co, c, k] += sum_{b, to} dy[b, co, to] x[b, g cg + c, to stride + k - pad] : one CTA per (co, c), fixed summation order
__global__ void __launch_bounds__(256) gconv_wgrad_kernel(const GConvParams p) {
    __shared__ float red[8][GC_KMAX];
    const int cg = p.Cin / p.groups, og = p.Cout / p.groups;
    const int co = blockIdx.x, c = blockIdx.y, g = co / og;
    float acc[GC_KMAX];
    for (int k = 0; k < p.K; ++k) acc[k] = 0.f;
    const int P = p.B * p.Tout;
    for (int q = threadIdx.x; q < P; q += 256) {
        const int b = q / p.Tout, to = q - b * p.Tout;
        const float d = p.dy[((size_t)b * p.Cout + co) * p.Tout + to];
        const float* xr = p.x + ((size_t)b * p.Cin + g * cg + c) * p.Tin;
        const int t0 = to * p.stride - p.pad;
        for (int k = 0; k < p.K; ++k) {
            const int ti = t0 + k;
            if (ti >= 0 && ti < p.Tin) acc[k] = fmaf(d, xr[ti], acc[k]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < p.K; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += 256) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w][k];
        p.dw[((size_t)co * cg + c) * p.K + k] += s;
    }
}

static int gconv_setup(GConvParams& p, int B, int Cin, int Tin, int Cout, int K, int stride, int pad, int groups) {
    TTTS_CHECK_ARG(B > 0 && Cin > 0 && Tin > 0 && Cout > 0 && K > 0 && K <= GC_KMAX && stride > 0 && pad >= 0 && groups > 0, "grouped conv: bad shape");
    TTTS_CHECK_ARG(Cin % groups == 0 && Cout % groups == 0, "grouped conv: channels (%d, %d) not divisible by groups %d", Cin, Cout, groups);
    const int Tout = (Tin + 2 * pad - K) / stride + 1;
    TTTS_CHECK_ARG(Tout > 0, "grouped conv: empty output");
    p.B = B; p.Cin = Cin; p.Tin = Tin; p.Cout = Cout; p.Tout = Tout; p.K = K; p.stride = stride; p.pad = pad; p.groups = groups;
    return TTTS_OK;
}
static inline unsigned gc_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(b > cap ? cap : (b ? b : 1));
}

}  // namespace ttts

using namespace ttts;

/* x [B,Cin,Tin], w [Cout, Cin/groups, K], y [B,Cout,Tout] ; dilation 1 */
extern "C" int ttts_gconv1d(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                            int32_t stride, int32_t pad, int32_t groups, void* stream) {
    GConvParams p = {};
    TTTS_RUN(gconv_setup(p, B, Cin, Tin, Cout, K, stride, pad, groups));
    TTTS_CHECK_ARG(x && w && y, "grouped conv: null pointer");
    p.x = x; p.w = w; p.bias = bias; p.y = y;
    TTTS_CUDA(launch_plain(gconv_fwd_kernel, dim3(gc_blocks((size_t)B * Cout * p.Tout)), dim3(256), 0, (cudaStream_t)stream, p));
    TTTS_LAUNCH_CHECK("gconv_fwd");
    return TTTS_OK;
}
/* dx written (may be NULL to skip) ; dw ACCUMULATES (zero it first; may be NULL to skip) */
extern "C" int ttts_gconv1d_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout,
                                int32_t K, int32_t stride, int32_t pad, int32_t groups, void* stream) {
    GConvParams p = {};
    TTTS_RUN(gconv_setup(p, B, Cin, Tin, Cout, K, stride, pad, groups));
    TTTS_CHECK_ARG(dy && x && w, "grouped conv backward: null pointer");
    TTTS_CHECK_ARG(Cout <= 65535 * 1 && Cin / groups <= 65535, "grouped conv backward: too many channels");
    p.dy = dy; p.x = x; p.w = w; p.dx = dx; p.dw = dw;
    if (dx) {
        TTTS_CUDA(launch_plain(gconv_dgrad_kernel, dim3(gc_blocks((size_t)B * Cin * Tin)), dim3(256), 0, (cudaStream_t)stream, p));
        TTTS_LAUNCH_CHECK("gconv_dgrad");
    }
    if (dw) {
        TTTS_CUDA(launch_plain(gconv_wgrad_kernel, dim3(Cout, Cin / groups), dim3(256), 0, (cudaStream_t)stream, p));
        TTTS_LAUNCH_CHECK("gconv_wgrad");
    }
    return TTTS_OK;
}
