// NOT YET RUN ON HARDWARE (written after the round's GPU budget was spent; validated on the CPU emulation of this source, tests/emu,
// against the per-op contract tests/ref_kernels.py).  Next scope row: the element-wise / small kernels that the TRAINING
// graph of the VQ-VAE encode half needs besides the convolutions (ttts_b200/vqvae/train_encoder.py): unfused forward of the gates and
// activations (their inputs must be kept for the backward), and the backward of weight norm, GLU, Mish, the WN gate, the anti-aliased
// SnakeBeta, the small masked attention of MelStyleEncoder, the masked mean and the posterior sample.
// All tensors fp32, [B, C, T] channel-major like the forward kernels (conv1d.cu).  Bandwidth- or latency-trivial: one pass, coalesced.
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

TTTS_DEVICE float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ---- element-wise -------------------------------------------------------------------------------------------
__global__ void ew_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = a[i] + b[i];
}
__global__ void ew_scale_kernel(const float* __restrict__ a, float s, float* __restrict__ o, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = a[i] * s;
}
// o[b, c, t] = a[b, c, t] * mask[b, t]
__global__ void ew_mul_mask_kernel(const float* __restrict__ a, const float* __restrict__ mask, float* __restrict__ o, int C, int T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)C * T);
        const int t = (int)(i % T);
        o[i] = a[i] * mask[b * T + t];
    }
}
// GLU (modules.py:560-566 Conv1dGLU): y[b, c, t] = a * sigmoid(g), raw = [a | g] over 2C channels.  dir = 0 forward, 1 backward (o = d raw)
__global__ void glu_kernel(const float* __restrict__ raw, const float* __restrict__ dy, float* __restrict__ o, int C, int T, size_t n, int dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)C * T), r = i - b * C * T;
        const float a = raw[b * 2 * C * T + r], g = raw[b * 2 * C * T + (size_t)C * T + r];
        const float s = sigmoid_f(g);
        if (dir == 0) { o[i] = a * s; continue; }
        const float d = dy[i];
        o[b * 2 * C * T + r] = d * s;
        o[b * 2 * C * T + (size_t)C * T + r] = d * a * s * (1.f - s);
    }
}
// Mish: x * tanh(softplus(x)) ; derivative tanh(sp) + x * (1 - tanh(sp)^2) * sigmoid(x)
__global__ void mish_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ o, size_t n, int dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float sp = v > 20.f ? v : log1pf(expf(v));
        const float th = tanhf(sp);
        o[i] = dir == 0 ? v * th : dy[i] * (th + v * (1.f - th * th) * sigmoid_f(v));
    }
}
// leaky ReLU with an explicit slope (the Generator uses 0.1 between stages and torch's default 0.01 before conv_post, vq2.py:392,404)
__global__ void lrelu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ o, size_t n, float slope, int dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        o[i] = dir == 0 ? (v > 0.f ? v : slope * v) : dy[i] * (v > 0.f ? 1.f : slope);
    }
}
__global__ void tanh_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ o, size_t n, int dir) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float t = tanhf(x[i]);
        o[i] = dir == 0 ? t : dy[i] * (1.f - t * t);
    }
}
// o[b, c, t] = a[b, c, t] + v[c] (per_batch = 0: a bias) or + v[b, c] (per_batch = 1: a conditioning vector)
__global__ void add_bcast_kernel(const float* __restrict__ a, const float* __restrict__ v, float* __restrict__ o, int C, int T, size_t n, int per_batch) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bc = i / T;
        o[i] = a[i] + v[per_batch ? bc : bc % C];
    }
}
// o[b, c] = sum_t a[b, c, t] : one warp per row, fixed order
__global__ void __launch_bounds__(32) sum_t_kernel(const float* __restrict__ a, float* __restrict__ o, int T) {
    const size_t row = blockIdx.x;
    float s = 0.f;
    for (int t = threadIdx.x; t < T; t += 32) s += a[row * T + t];
    s = warp_sum(s);
    if (threadIdx.x == 0) o[row] = s;
}
// posterior sample backward: z = (m + eps e^logs) mask  ->  dm = dz mask ; dlogs = dz mask eps e^logs          (vq2.py:742-744)
__global__ void posterior_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ stats, const float* __restrict__ eps,
                                     const float* __restrict__ mask, float* __restrict__ dstats, int C, int T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)C * T), r = i - b * C * T;
        const int t = (int)(r % T);
        const float d = dz[i] * (mask ? mask[b * T + t] : 1.f);
        dstats[b * 2 * C * T + r] = d;
        dstats[b * 2 * C * T + (size_t)C * T + r] = eps ? d * eps[i] * expf(stats[b * 2 * C * T + (size_t)C * T + r]) : 0.f;
    }
}
// masked mean backward: dx[b, c, t] = dy[b, c] / len[b] for t < len[b], else 0                                  (modules.py:757-763)
__global__ void masked_mean_bwd_kernel(const float* __restrict__ dy, const int64_t* __restrict__ lens, float* __restrict__ dx, int C, int T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bc = i / T;
        const int t = (int)(i - bc * T);
        const size_t b = bc / C;
        const int len = lens ? (int)min((int64_t)T, lens[b]) : T;
        dx[i] = t < len ? dy[bc] / (float)len : 0.f;
    }
}

// ---- weight norm backward: w = g v / ||v|| per output channel -> dg = <dw, v> / ||v|| ; dv = (g / ||v||) (dw - v <dw, v> / ||v||^2) ----
__global__ void __launch_bounds__(128) weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                                                              float* __restrict__ dv, float* __restrict__ dg, int n) {
    __shared__ float sm[2][4];
    const int co = blockIdx.x;
    const float* vr = v + (size_t)co * n;
    const float* dr = dw + (size_t)co * n;
    float ss = 0.f, dot = 0.f;
    for (int i = threadIdx.x; i < n; i += 128) { ss = fmaf(vr[i], vr[i], ss); dot = fmaf(dr[i], vr[i], dot); }
    ss = warp_sum(ss); dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = ss; sm[1][threadIdx.x >> 5] = dot; }
    __syncthreads();
    ss = (sm[0][0] + sm[0][1]) + (sm[0][2] + sm[0][3]);
    dot = (sm[1][0] + sm[1][1]) + (sm[1][2] + sm[1][3]);
    const float inv = 1.f / sqrtf(ss);
    if (threadIdx.x == 0) dg[co] = dot * inv;
    const float sc = g[co] * inv, proj = dot / ss;
    for (int i = threadIdx.x; i < n; i += 128) dv[(size_t)co * n + i] = sc * (dr[i] - vr[i] * proj);
}

// ---- WN gate (modules.py:195-201): y[b, c, t] = tanh(a + cond[b, c]) * sigmoid(g + cond[b, H + c]), raw = [a | g] over 2H channels ----
// one warp per (b, c) row; backward also reduces the conditioning gradient over t
__global__ void __launch_bounds__(32) gate_kernel(const float* __restrict__ raw, const float* __restrict__ cond, const float* __restrict__ dy,
                                                  float* __restrict__ o, float* __restrict__ dcond, int H, int T, int dir) {
    const int c = blockIdx.x, b = blockIdx.y;
    const float ca = cond ? cond[(size_t)b * 2 * H + c] : 0.f, cg = cond ? cond[(size_t)b * 2 * H + H + c] : 0.f;
    const float* ar = raw + ((size_t)b * 2 * H + c) * T;
    const float* gr = raw + ((size_t)b * 2 * H + H + c) * T;
    float sa = 0.f, sg = 0.f;
    for (int t = threadIdx.x; t < T; t += 32) {
        const float th = tanhf(ar[t] + ca), s = sigmoid_f(gr[t] + cg);
        if (dir == 0) { o[((size_t)b * H + c) * T + t] = th * s; continue; }
        const float d = dy[((size_t)b * H + c) * T + t];
        const float da = d * s * (1.f - th * th), dg = d * th * s * (1.f - s);
        o[((size_t)b * 2 * H + c) * T + t] = da;
        o[((size_t)b * 2 * H + H + c) * T + t] = dg;
        sa += da; sg += dg;
    }
    if (dir != 0 && dcond) {
        sa = warp_sum(sa); sg = warp_sum(sg);
        if (threadIdx.x == 0) { dcond[(size_t)b * 2 * H + c] = sa; dcond[(size_t)b * 2 * H + H + c] = sg; }
    }
}

// ---- Activation1d(SnakeBeta) backward (forward: conv1d.cu snake_aa_kernel; alias_free_torch/act.py:8-28, activations.py:62-119) ----
//   xp = replicate-pad(x, 5, 5) ; s[i] = 2 sum_k xp[(i + 15 - k) / 2] f[k] ((i + 15 - k) even) ; a[i] = s + sin^2(alpha s) / (beta + 1e-9)
//   up = replicate-pad(a, 5, 6) ; y[t] = sum_k up[2 t + k] f[k]
// one block per (b, c) row; d log_alpha / d log_beta accumulate over b with atomic adds (outputs zeroed by the caller)
__global__ void __launch_bounds__(64) snake_aa_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ log_alpha,
                                                          const float* __restrict__ log_beta, const float* __restrict__ filt, float* __restrict__ dx,
                                                          float* __restrict__ dla, float* __restrict__ dlb, int C, int T) {
    TTTS_DYN_SMEM(float, sn);
    __shared__ float red[2][2];
    const int row = blockIdx.x, c = row % C;
    const float* xr = x + (size_t)row * T;
    const float* dyr = dy + (size_t)row * T;
    float* xp = sn;                        // [T + 10]
    float* dup = xp + (T + 10);            // [2T + 11] gradient of the padded up-sampled signal
    float* ds = dup + (2 * T + 11);        // [2T]
    float* dxp = ds + 2 * T;               // [T + 10]
    const float alpha = expf(log_alpha[c]), beta = expf(log_beta[c]);
    const float ib = 1.f / (beta + 1e-9f);
    for (int i = threadIdx.x; i < T + 10; i += 64) xp[i] = xr[min(max(i - 5, 0), T - 1)];
    // down filter transposed: dup[j] = sum_{t, k : 2 t + k = j} dy[t] f[k]
    for (int j = threadIdx.x; j < 2 * T + 11; j += 64) {
        float s = 0.f;
        for (int k = 0; k < 12; ++k) {
            const int m = j - k;
            if (m >= 0 && (m & 1) == 0 && (m >> 1) < T) s += dyr[m >> 1] * filt[k];
        }
        dup[j] = s;
    }
    __syncthreads();
    // fold the replicate padding (5 left, 6 right) into the first / last sample, then through the activation
    float ga = 0.f, gb = 0.f;
    for (int i = threadIdx.x; i < 2 * T; i += 64) {
        float da = dup[5 + i];
        if (i == 0) for (int j = 0; j < 5; ++j) da += dup[j];
        if (i == 2 * T - 1) for (int j = 0; j < 6; ++j) da += dup[5 + 2 * T + j];
        const int n = i + 15;
        float s = 0.f;
        for (int k = 0; k < 12; ++k) {
            const int m = n - k;
            if (m >= 0 && (m & 1) == 0 && (m >> 1) < T + 10) s += xp[m >> 1] * filt[k];
        }
        s *= 2.f;
        const float sv = sinf(alpha * s), s2 = sinf(2.f * alpha * s);
        ds[i] = da * (1.f + ib * alpha * s2);
        ga += da * ib * s2 * s * alpha;                  // d/d log_alpha = d/d alpha * alpha
        gb -= da * sv * sv * ib * ib * beta;             // d/d log_beta  = d/d beta  * beta
    }
    ga = warp_sum(ga); gb = warp_sum(gb);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ga; red[1][threadIdx.x >> 5] = gb; }
    __syncthreads();
    if (threadIdx.x == 0) { atomicAdd(dla + c, red[0][0] + red[0][1]); atomicAdd(dlb + c, red[1][0] + red[1][1]); }
    // up filter transposed: dxp[m] = 2 sum_{i, k : (i + 15 - k) = 2 m} ds[i] f[k]
    for (int m = threadIdx.x; m < T + 10; m += 64) {
        float s = 0.f;
        for (int k = 0; k < 12; ++k) {
            const int i = 2 * m + k - 15;
            if (i >= 0 && i < 2 * T) s += ds[i] * filt[k];
        }
        dxp[m] = 2.f * s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += 64) {
        float s = dxp[t + 5];
        if (t == 0) for (int j = 0; j < 5; ++j) s += dxp[j];
        if (t == T - 1) for (int j = 0; j < 5; ++j) s += dxp[T + 5 + j];
        dx[(size_t)row * T + t] = s;
    }
}

// ---- small masked attention backward (forward: conv1d.cu mha_small_kernel; MelStyleEncoder's MultiHeadAttention, modules.py:640-683) ----
// one block per (head, batch), thread = query row tq.  P is recomputed; dV[j, tk] = sum_tq P[tq, tk] dO[j, tq] ;
// dS[tq, tk] = P (dP - sum_tk' P dP) with dP[tq, tk] = sum_j dO[j, tq] V[j, tk] ; dQ[j, tq] = sum_tk dS K[j, tk] / temp ; dK[j, tk] = sum_tq dS Q[j, tq] / temp
__global__ void __launch_bounds__(64) mha_small_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ q, const float* __restrict__ k,
                                                           const float* __restrict__ v, const int64_t* __restrict__ lens, float* __restrict__ dq,
                                                           float* __restrict__ dk_out, float* __restrict__ dv, int C, int T, int dk, float inv_temp) {
    TTTS_DYN_SMEM(float, ms);
    const int b = blockIdx.y, h = blockIdx.x;
    float* sq = ms;                 // [dk][T]
    float* sk = sq + dk * T;
    float* sv = sk + dk * T;
    float* sdo = sv + dk * T;
    float* sP = sdo + dk * T;       // [T][T + 1] probabilities, then dS
    const int LP = T + 1;
    const size_t base = ((size_t)b * C + (size_t)h * dk) * T;
    for (int i = threadIdx.x; i < dk * T; i += 64) { sq[i] = q[base + i]; sk[i] = k[base + i]; sv[i] = v[base + i]; sdo[i] = dout[base + i]; }
    __syncthreads();
    const int len = lens ? (int)min((int64_t)T, lens[b]) : T;
    const int tq = threadIdx.x;
    if (tq < T) {
        float mx = -INFINITY;
        for (int tk = 0; tk < T; ++tk) {
            float a = 0.f;
            for (int j = 0; j < dk; ++j) a = fmaf(sq[j * T + tq], sk[j * T + tk], a);
            a = (tk < len) ? a * inv_temp : -INFINITY;
            sP[tq * LP + tk] = a;
            mx = fmaxf(mx, a);
        }
        float sum = 0.f;
        for (int tk = 0; tk < T; ++tk) { const float e = expf(sP[tq * LP + tk] - mx); sP[tq * LP + tk] = e; sum += e; }
        const float inv = 1.f / sum;
        for (int tk = 0; tk < T; ++tk) sP[tq * LP + tk] *= inv;
    }
    __syncthreads();
    // dV (thread = key column tk here): needs all rows of P
    if (tq < T) {
        const int tk = tq;
        for (int j = 0; j < dk; ++j) {
            float a = 0.f;
            for (int t2 = 0; t2 < T; ++t2) a = fmaf(sP[t2 * LP + tk], sdo[j * T + t2], a);
            dv[base + (size_t)j * T + tk] = a;
        }
    }
    __syncthreads();
    // dS in place (row tq): dS = P (dP - <P, dP>)
    if (tq < T) {
        float dprow[64];
        float dot = 0.f;
        for (int tk = 0; tk < T; ++tk) {
            float dp = 0.f;
            for (int j = 0; j < dk; ++j) dp = fmaf(sdo[j * T + tq], sv[j * T + tk], dp);
            dprow[tk] = dp;
            dot = fmaf(sP[tq * LP + tk], dp, dot);
        }
        for (int tk = 0; tk < T; ++tk) sP[tq * LP + tk] *= dprow[tk] - dot;
        // dQ
        for (int j = 0; j < dk; ++j) {
            float a = 0.f;
            for (int tk = 0; tk < T; ++tk) a = fmaf(sP[tq * LP + tk], sk[j * T + tk], a);
            dq[base + (size_t)j * T + tq] = a * inv_temp;
        }
    }
    __syncthreads();
    // dK (thread = key column tk)
    if (tq < T) {
        const int tk = tq;
        for (int j = 0; j < dk; ++j) {
            float a = 0.f;
            for (int t2 = 0; t2 < T; ++t2) a = fmaf(sP[t2 * LP + tk], sq[j * T + t2], a);
            dk_out[base + (size_t)j * T + tk] = a * inv_temp;
        }
    }
}

static inline unsigned ew_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (unsigned)(b > cap ? cap : (b ? b : 1));
}

}  // namespace ttts

using namespace ttts;

#define TTTS_API extern "C"

TTTS_API int ttts_ew_add(const float* a, const float* b, float* o, int64_t n, void* stream) {
    TTTS_CHECK_ARG(a && b && o && n > 0, "ew_add: bad args");
    TTTS_CUDA(launch_plain(ew_add_kernel, dim3(ew_blocks((size_t)n)), dim3(256), 0, (cudaStream_t)stream, a, b, o, (size_t)n));
    TTTS_LAUNCH_CHECK("ew_add");
    return TTTS_OK;
}
TTTS_API int ttts_ew_scale(const float* a, float s, float* o, int64_t n, void* stream) {
    TTTS_CHECK_ARG(a && o && n > 0, "ew_scale: bad args");
    TTTS_CUDA(launch_plain(ew_scale_kernel, dim3(ew_blocks((size_t)n)), dim3(256), 0, (cudaStream_t)stream, a, s, o, (size_t)n));
    TTTS_LAUNCH_CHECK("ew_scale");
    return TTTS_OK;
}
TTTS_API int ttts_ew_mul_mask(const float* a, const float* mask, float* o, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(a && mask && o && B > 0 && C > 0 && T > 0, "ew_mul_mask: bad args");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(ew_mul_mask_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, a, mask, o, C, T, n));
    TTTS_LAUNCH_CHECK("ew_mul_mask");
    return TTTS_OK;
}
/* raw [B, 2C, T] ; forward: y [B, C, T] ; backward: draw [B, 2C, T] from dy [B, C, T] */
TTTS_API int ttts_glu(const float* raw, const float* dy, float* out, int32_t B, int32_t C, int32_t T, int32_t backward, void* stream) {
    TTTS_CHECK_ARG(raw && out && (!backward || dy) && B > 0 && C > 0 && T > 0, "glu: bad args");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(glu_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, raw, dy, out, C, T, n, backward));
    TTTS_LAUNCH_CHECK("glu");
    return TTTS_OK;
}
TTTS_API int ttts_mish(const float* x, const float* dy, float* out, int64_t n, int32_t backward, void* stream) {
    TTTS_CHECK_ARG(x && out && (!backward || dy) && n > 0, "mish: bad args");
    TTTS_CUDA(launch_plain(mish_kernel, dim3(ew_blocks((size_t)n)), dim3(256), 0, (cudaStream_t)stream, x, dy, out, (size_t)n, backward));
    TTTS_LAUNCH_CHECK("mish");
    return TTTS_OK;
}
/* raw [B, 2H, T], cond [B, 2H] or NULL ; forward: y [B, H, T] ; backward: draw [B, 2H, T], dcond [B, 2H] (may be NULL) from dy [B, H, T] */
TTTS_API int ttts_wn_gate(const float* raw, const float* cond, const float* dy, float* out, float* dcond, int32_t B, int32_t H, int32_t T,
                          int32_t backward, void* stream) {
    TTTS_CHECK_ARG(raw && out && (!backward || dy) && B > 0 && B <= 65535 && H > 0 && T > 0, "wn_gate: bad args");
    TTTS_CUDA(launch_plain(gate_kernel, dim3(H, B), dim3(32), 0, (cudaStream_t)stream, raw, cond, dy, out, dcond, H, T, backward));
    TTTS_LAUNCH_CHECK("wn_gate");
    return TTTS_OK;
}
TTTS_API int ttts_lrelu(const float* x, const float* dy, float* out, int64_t n, float slope, int32_t backward, void* stream) {
    TTTS_CHECK_ARG(x && out && (!backward || dy) && n > 0, "lrelu: bad args");
    TTTS_CUDA(launch_plain(lrelu_kernel, dim3(ew_blocks((size_t)n)), dim3(256), 0, (cudaStream_t)stream, x, dy, out, (size_t)n, slope, backward));
    TTTS_LAUNCH_CHECK("lrelu");
    return TTTS_OK;
}
TTTS_API int ttts_tanh(const float* x, const float* dy, float* out, int64_t n, int32_t backward, void* stream) {
    TTTS_CHECK_ARG(x && out && (!backward || dy) && n > 0, "tanh: bad args");
    TTTS_CUDA(launch_plain(tanh_kernel, dim3(ew_blocks((size_t)n)), dim3(256), 0, (cudaStream_t)stream, x, dy, out, (size_t)n, backward));
    TTTS_LAUNCH_CHECK("tanh");
    return TTTS_OK;
}
TTTS_API int ttts_add_bcast(const float* a, const float* v, float* out, int32_t B, int32_t C, int32_t T, int32_t per_batch, void* stream) {
    TTTS_CHECK_ARG(a && v && out && B > 0 && C > 0 && T > 0, "add_bcast: bad args");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(add_bcast_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, a, v, out, C, T, n, per_batch));
    TTTS_LAUNCH_CHECK("add_bcast");
    return TTTS_OK;
}
TTTS_API int ttts_sum_t(const float* a, float* out, int32_t rows, int32_t T, void* stream) {
    TTTS_CHECK_ARG(a && out && rows > 0 && T > 0, "sum_t: bad args");
    TTTS_CUDA(launch_plain(sum_t_kernel, dim3(rows), dim3(32), 0, (cudaStream_t)stream, a, out, T));
    TTTS_LAUNCH_CHECK("sum_t");
    return TTTS_OK;
}
TTTS_API int ttts_weight_norm_bwd(const float* dw, const float* v, const float* g, float* dv, float* dg, int32_t Cout, int32_t n_per_out, void* stream) {
    TTTS_CHECK_ARG(dw && v && g && dv && dg && Cout > 0 && n_per_out > 0, "weight_norm_bwd: bad args");
    TTTS_CUDA(launch_plain(weight_norm_bwd_kernel, dim3(Cout), dim3(128), 0, (cudaStream_t)stream, dw, v, g, dv, dg, n_per_out));
    TTTS_LAUNCH_CHECK("weight_norm_bwd");
    return TTTS_OK;
}
/* dla / dlb [C] ACCUMULATE (zero them first) */
TTTS_API int ttts_snake_aa_bwd(const float* dy, const float* x, const float* log_alpha, const float* log_beta, const float* filt12, float* dx,
                               float* dla, float* dlb, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(dy && x && log_alpha && log_beta && filt12 && dx && dla && dlb && B > 0 && C > 0 && T > 0, "snake_aa_bwd: bad args");
    const size_t smem = (size_t)((T + 10) * 2 + (2 * T + 11) + 2 * T) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 48 * 1024, "snake_aa_bwd: T too large for shared memory (%d)", T);
    TTTS_CUDA(launch_plain(snake_aa_bwd_kernel, dim3(B * C), dim3(64), smem, (cudaStream_t)stream, dy, x, log_alpha, log_beta, filt12, dx, dla, dlb, C, T));
    TTTS_LAUNCH_CHECK("snake_aa_bwd");
    return TTTS_OK;
}
TTTS_API int ttts_mha_small_bwd(const float* dout, const float* q, const float* k, const float* v, const int64_t* lens, float* dq, float* dk, float* dv,
                                int32_t B, int32_t C, int32_t T, int32_t heads, float temperature, void* stream) {
    TTTS_CHECK_ARG(dout && q && k && v && dq && dk && dv && B > 0 && C > 0 && heads > 0 && C % heads == 0, "mha_small_bwd: bad args");
    TTTS_CHECK_ARG(T >= 1 && T <= 64, "mha_small_bwd: T must be <= 64 (got %d)", T);
    const int d = C / heads;
    const size_t smem = ((size_t)4 * d * T + (size_t)T * (T + 1)) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "mha_small_bwd: head too large");
#ifndef TTTS_HOST_EMU
    static size_t attr = 48 * 1024;
    if (smem > attr) { TTTS_CUDA(cudaFuncSetAttribute(mha_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
#endif
    TTTS_CUDA(launch_plain(mha_small_bwd_kernel, dim3(heads, B), dim3(64), smem, (cudaStream_t)stream, dout, q, k, v, lens, dq, dk, dv, C, T, d,
                           1.0f / temperature));
    TTTS_LAUNCH_CHECK("mha_small_bwd");
    return TTTS_OK;
}
TTTS_API int ttts_masked_mean_bwd(const float* dy, const int64_t* lens, float* dx, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(dy && dx && B > 0 && C > 0 && T > 0, "masked_mean_bwd: bad args");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(masked_mean_bwd_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, dy, lens, dx, C, T, n));
    TTTS_LAUNCH_CHECK("masked_mean_bwd");
    return TTTS_OK;
}
TTTS_API int ttts_posterior_sample_bwd(const float* dz, const float* stats, const float* eps, const float* mask, float* dstats, int32_t B, int32_t C,
                                       int32_t T, void* stream) {
    TTTS_CHECK_ARG(dz && stats && dstats && B > 0 && C > 0 && T > 0, "posterior_sample_bwd: bad args");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(posterior_bwd_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, dz, stats, eps, mask, dstats, C, T, n));
    TTTS_LAUNCH_CHECK("posterior_sample_bwd");
    return TTTS_OK;
}
