// NOT YET RUN ON HARDWARE (validated on the CPU emulation of this source against tests/ref_kernels.py).  Next scope row:
// the two ops the prior encoder enc_p_2 (TextEncoder + MRTE, ttts/vqvae/vq2.py:17-164) adds to the training tape:
//   * multi-head attention over short sequences, channel-major [B, C, T], with an optional windowed relative-position term
//     (attentions.py:231-363 restated in closed form: scores[i,j] += [|j-i| <= w] q_i . Ek[j-i+w], out[i] += sum_{|j-i| <= w} p[i,j] Ev[j-i+w]),
//     different query / key lengths (MRTE's cross attention, vc_utils.py:568-640) and the reference's -1e4 masking;
//   * LayerNorm over the channel axis (modules.py:20-32).
// Correctness first: one CTA per (batch, head), the probability matrix in shared memory, operands read through the cache.
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

struct AttnParams {
    const float *q, *k, *v, *emb_k, *emb_v, *dout;      // emb_* [2w+1, dk] or null
    const int64_t *q_len, *k_len;
    float *out, *dq, *dk, *dv, *demb_k, *demb_v;
    int C, Tq, Tk, heads, dk_, win;
};

// probabilities of one (b, h) into sP [Tq][Tk + 1]; thread = query rows i, i + blockDim, ...
TTTS_DEVICE void attn_probs(const AttnParams& p, int b, int h, float* sP) {
    const int dk = p.dk_, LP = p.Tk + 1;
    const float sc = rsqrtf((float)dk);
    const int ql = (int)min((int64_t)p.Tq, p.q_len[b]), kl = (int)min((int64_t)p.Tk, p.k_len[b]);
    const float* qb = p.q + ((size_t)b * p.C + (size_t)h * dk) * p.Tq;
    const float* kb = p.k + ((size_t)b * p.C + (size_t)h * dk) * p.Tk;
    for (int i = threadIdx.x; i < p.Tq; i += blockDim.x) {
        float mx = -INFINITY;
        for (int j = 0; j < p.Tk; ++j) {
            float s = 0.f;
            const int r = j - i + p.win;
            const bool rel = p.emb_k && r >= 0 && r <= 2 * p.win;
            for (int d = 0; d < dk; ++d) {
                const float kv = kb[(size_t)d * p.Tk + j] + (rel ? p.emb_k[r * dk + d] : 0.f);
                s = fmaf(qb[(size_t)d * p.Tq + i] * sc, kv, s);
            }
            if (!(i < ql && j < kl)) s = -1e4f;
            sP[i * LP + j] = s;
            mx = fmaxf(mx, s);
        }
        float sum = 0.f;
        for (int j = 0; j < p.Tk; ++j) { const float e = expf(sP[i * LP + j] - mx); sP[i * LP + j] = e; sum += e; }
        const float inv = 1.f / sum;
        for (int j = 0; j < p.Tk; ++j) sP[i * LP + j] *= inv;
    }
}

__global__ void __launch_bounds__(128) attn_small_fwd_kernel(const AttnParams p) {
    TTTS_DYN_SMEM(float, sP);
    const int h = blockIdx.x, b = blockIdx.y, dk = p.dk_, LP = p.Tk + 1;
    attn_probs(p, b, h, sP);
    __syncthreads();
    const float* vb = p.v + ((size_t)b * p.C + (size_t)h * dk) * p.Tk;
    float* ob = p.out + ((size_t)b * p.C + (size_t)h * dk) * p.Tq;
    for (int i = threadIdx.x; i < p.Tq; i += blockDim.x) {
        for (int d = 0; d < dk; ++d) {
            float a = 0.f;
            for (int j = 0; j < p.Tk; ++j) {
                const int r = j - i + p.win;
                const float vv = vb[(size_t)d * p.Tk + j] + ((p.emb_v && r >= 0 && r <= 2 * p.win) ? p.emb_v[r * dk + d] : 0.f);
                a = fmaf(sP[i * LP + j], vv, a);
            }
            ob[(size_t)d * p.Tq + i] = a;
        }
    }
}

// dq, dk, dv written; demb_k / demb_v ACCUMULATE (shared by heads and batch): zero them first
__global__ void __launch_bounds__(128) attn_small_bwd_kernel(const AttnParams p) {
    TTTS_DYN_SMEM(float, sm);
    const int h = blockIdx.x, b = blockIdx.y, dk = p.dk_, LP = p.Tk + 1;
    float* sP = sm;                          // [Tq][LP] probabilities
    float* sS = sm + (size_t)p.Tq * LP;      // [Tq][LP] dS
    attn_probs(p, b, h, sP);
    __syncthreads();
    const float sc = rsqrtf((float)dk);
    const int ql = (int)min((int64_t)p.Tq, p.q_len[b]), kl = (int)min((int64_t)p.Tk, p.k_len[b]);
    const size_t qoff = ((size_t)b * p.C + (size_t)h * dk) * p.Tq, koff = ((size_t)b * p.C + (size_t)h * dk) * p.Tk;
    const float *qb = p.q + qoff, *kb = p.k + koff, *vb = p.v + koff, *dob = p.dout + qoff;
    // dS rows
    for (int i = threadIdx.x; i < p.Tq; i += blockDim.x) {
        float dot = 0.f;
        for (int j = 0; j < p.Tk; ++j) {
            const int r = j - i + p.win;
            const bool rel = p.emb_v && r >= 0 && r <= 2 * p.win;
            float dp = 0.f;
            for (int d = 0; d < dk; ++d) dp = fmaf(dob[(size_t)d * p.Tq + i], vb[(size_t)d * p.Tk + j] + (rel ? p.emb_v[r * dk + d] : 0.f), dp);
            sS[i * LP + j] = dp;
            dot = fmaf(sP[i * LP + j], dp, dot);
        }
        for (int j = 0; j < p.Tk; ++j) {
            const float ds = sP[i * LP + j] * (sS[i * LP + j] - dot);
            sS[i * LP + j] = (i < ql && j < kl) ? ds : 0.f;              // masked_fill: no gradient reaches the masked scores
        }
        // dq row
        for (int d = 0; d < dk; ++d) {
            float a = 0.f;
            for (int j = 0; j < p.Tk; ++j) {
                const int r = j - i + p.win;
                const float kv = kb[(size_t)d * p.Tk + j] + ((p.emb_k && r >= 0 && r <= 2 * p.win) ? p.emb_k[r * dk + d] : 0.f);
                a = fmaf(sS[i * LP + j], kv, a);
            }
            p.dq[qoff + (size_t)d * p.Tq + i] = a * sc;
        }
    }
    __syncthreads();
    // dk, dv columns (thread = key j)
    for (int j = threadIdx.x; j < p.Tk; j += blockDim.x) {
        for (int d = 0; d < dk; ++d) {
            float ak = 0.f, av = 0.f;
            for (int i = 0; i < p.Tq; ++i) {
                ak = fmaf(sS[i * LP + j], qb[(size_t)d * p.Tq + i], ak);
                av = fmaf(sP[i * LP + j], dob[(size_t)d * p.Tq + i], av);
            }
            p.dk[koff + (size_t)d * p.Tk + j] = ak * sc;
            p.dv[koff + (size_t)d * p.Tk + j] = av;
        }
    }
    // relative-position tables: thread = (r, d)
    if (p.emb_k) {
        for (int t = threadIdx.x; t < (2 * p.win + 1) * dk; t += blockDim.x) {
            const int r = t / dk, d = t - r * dk;
            float ak = 0.f, av = 0.f;
            for (int i = 0; i < p.Tq; ++i) {
                const int j = i + r - p.win;
                if (j < 0 || j >= p.Tk) continue;
                ak = fmaf(sS[i * LP + j], qb[(size_t)d * p.Tq + i], ak);
                av = fmaf(sP[i * LP + j], dob[(size_t)d * p.Tq + i], av);
            }
            atomicAdd(p.demb_k + t, ak * sc);
            atomicAdd(p.demb_v + t, av);
        }
    }
}

// ---- LayerNorm over channels of [B, C, T]: CTA = 32 time steps x 8 channel lanes ----
constexpr float kLncEps = 1e-5f;
// stats[(b T + t) * 2] = mean, rstd
__global__ void __launch_bounds__(256) lnc_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, int C, int T) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + tx, b = blockIdx.y;
    const float* xb = x + (size_t)b * C * T;
    float s = 0.f;
    if (t < T) for (int c = cy; c < C; c += 8) s += xb[(size_t)c * T + t];
    red[cy][tx] = s;
    __syncthreads();
    float mean = 0.f;
    for (int w = 0; w < 8; ++w) mean += red[w][tx];
    mean /= (float)C;
    __syncthreads();
    float q = 0.f;
    if (t < T) for (int c = cy; c < C; c += 8) { const float d = xb[(size_t)c * T + t] - mean; q = fmaf(d, d, q); }
    red[cy][tx] = q;
    __syncthreads();
    if (cy == 0 && t < T) {
        float var = 0.f;
        for (int w = 0; w < 8; ++w) var += red[w][tx];
        stats[((size_t)b * T + t) * 2] = mean;
        stats[((size_t)b * T + t) * 2 + 1] = rsqrtf(var / (float)C + kLncEps);
    }
}
__global__ void lnc_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ y, int C, int T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)C * T), r = i - b * C * T;
        const int c = (int)(r / T), t = (int)(r - (size_t)c * T);
        const float mean = stats[(b * T + t) * 2], rstd = stats[(b * T + t) * 2 + 1];
        y[i] = (x[i] - mean) * rstd * gamma[c] + beta[c];
    }
}
// per column: m1 = mean_c(gamma dy), m2 = mean_c(gamma dy xhat) -> colred[(b T + t) * 2]
__global__ void __launch_bounds__(256) lnc_bwd_col_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma, float* __restrict__ colred, int C, int T) {
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + tx, b = blockIdx.y;
    float a1 = 0.f, a2 = 0.f;
    if (t < T) {
        const float mean = stats[((size_t)b * T + t) * 2], rstd = stats[((size_t)b * T + t) * 2 + 1];
        for (int c = cy; c < C; c += 8) {
            const size_t o = ((size_t)b * C + c) * T + t;
            const float g = gamma[c] * dy[o];
            a1 += g;
            a2 = fmaf(g, (x[o] - mean) * rstd, a2);
        }
    }
    red[0][cy][tx] = a1; red[1][cy][tx] = a2;
    __syncthreads();
    if (cy == 0 && t < T) {
        float s1 = 0.f, s2 = 0.f;
        for (int w = 0; w < 8; ++w) { s1 += red[0][w][tx]; s2 += red[1][w][tx]; }
        colred[((size_t)b * T + t) * 2] = s1 / (float)C;
        colred[((size_t)b * T + t) * 2 + 1] = s2 / (float)C;
    }
}
__global__ void lnc_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                                  const float* __restrict__ colred, const float* __restrict__ gamma, float* __restrict__ dx, int C, int T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)C * T), r = i - b * C * T;
        const int c = (int)(r / T), t = (int)(r - (size_t)c * T);
        const float mean = stats[(b * T + t) * 2], rstd = stats[(b * T + t) * 2 + 1];
        const float xhat = (x[i] - mean) * rstd;
        dx[i] = rstd * (gamma[c] * dy[i] - colred[(b * T + t) * 2] - xhat * colred[(b * T + t) * 2 + 1]);
    }
}
// dgamma[c] = sum_{b,t} dy xhat ; dbeta[c] = sum dy : one CTA per channel, fixed order
__global__ void __launch_bounds__(256) lnc_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int C, int T) {
    __shared__ float red[2][8];
    const int c = blockIdx.x;
    float g = 0.f, s = 0.f;
    for (int i = threadIdx.x; i < B * T; i += 256) {
        const int b = i / T, t = i - b * T;
        const size_t o = ((size_t)b * C + c) * T + t;
        const float d = dy[o];
        g = fmaf(d, (x[o] - stats[(size_t)i * 2]) * stats[(size_t)i * 2 + 1], g);
        s += d;
    }
    g = warp_sum(g); s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = g; red[1][threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, e = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; e += red[1][w]; }
        dgamma[c] = a; dbeta[c] = e;
    }
}

static int attn_setup(AttnParams& p, int B, int C, int Tq, int Tk, int heads, int win, size_t& smem, bool bwd) {
    TTTS_CHECK_ARG(B >= 1 && B <= 65535 && C >= 1 && heads >= 1 && C % heads == 0 && Tq >= 1 && Tk >= 1 && win >= 0, "attn_small: bad shape");
    TTTS_CHECK_ARG(p.q && p.k && p.v && p.q_len && p.k_len, "attn_small: null pointer");
    TTTS_CHECK_ARG((p.emb_k == nullptr) == (p.emb_v == nullptr), "attn_small: emb_k and emb_v go together");
    p.C = C; p.Tq = Tq; p.Tk = Tk; p.heads = heads; p.dk_ = C / heads; p.win = win;
    smem = (size_t)Tq * (Tk + 1) * sizeof(float) * (bwd ? 2 : 1);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "attn_small: %d x %d probabilities do not fit in shared memory", Tq, Tk);
    return TTTS_OK;
}
static inline unsigned te_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    return (unsigned)(b > cap ? cap : (b ? b : 1));
}

}  // namespace ttts

using namespace ttts;

/* out [B,C,Tq] = attention(q [B,C,Tq], k / v [B,C,Tk]) ; emb_k / emb_v [2 win + 1, C / heads] or NULL ; q_len / k_len [B] int64 */
extern "C" int ttts_attn_small(const float* q, const float* k, const float* v, const float* emb_k, const float* emb_v, const int64_t* q_len,
                               const int64_t* k_len, float* out, int32_t B, int32_t C, int32_t Tq, int32_t Tk, int32_t heads, int32_t win, void* stream) {
    AttnParams p = {};
    p.q = q; p.k = k; p.v = v; p.emb_k = emb_k; p.emb_v = emb_v; p.q_len = q_len; p.k_len = k_len; p.out = out;
    size_t smem;
    TTTS_RUN(attn_setup(p, B, C, Tq, Tk, heads, win, smem, false));
    TTTS_CHECK_ARG(out != nullptr, "attn_small: null output");
#ifndef TTTS_HOST_EMU
    static size_t attr = 48 * 1024;
    if (smem > attr) { TTTS_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
#endif
    TTTS_CUDA(launch_plain(attn_small_fwd_kernel, dim3(heads, B), dim3(128), smem, (cudaStream_t)stream, p));
    TTTS_LAUNCH_CHECK("attn_small_fwd");
    return TTTS_OK;
}
/* dq / dk / dv written ; demb_k / demb_v ACCUMULATE (zero them first) */
extern "C" int ttts_attn_small_bwd(const float* dout, const float* q, const float* k, const float* v, const float* emb_k, const float* emb_v,
                                   const int64_t* q_len, const int64_t* k_len, float* dq, float* dk, float* dv, float* demb_k, float* demb_v,
                                   int32_t B, int32_t C, int32_t Tq, int32_t Tk, int32_t heads, int32_t win, void* stream) {
    AttnParams p = {};
    p.q = q; p.k = k; p.v = v; p.emb_k = emb_k; p.emb_v = emb_v; p.q_len = q_len; p.k_len = k_len; p.dout = dout;
    p.dq = dq; p.dk = dk; p.dv = dv; p.demb_k = demb_k; p.demb_v = demb_v;
    size_t smem;
    TTTS_RUN(attn_setup(p, B, C, Tq, Tk, heads, win, smem, true));
    TTTS_CHECK_ARG(dout && dq && dk && dv && (!emb_k || (demb_k && demb_v)), "attn_small backward: null pointer");
#ifndef TTTS_HOST_EMU
    static size_t attr = 48 * 1024;
    if (smem > attr) { TTTS_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
#endif
    TTTS_CUDA(launch_plain(attn_small_bwd_kernel, dim3(heads, B), dim3(128), smem, (cudaStream_t)stream, p));
    TTTS_LAUNCH_CHECK("attn_small_bwd");
    return TTTS_OK;
}
/* y = LayerNorm_C(x) gamma + beta for x [B,C,T] ; stats: B*T*2 floats (kept for the backward) */
extern "C" int ttts_layernorm_c(const float* x, const float* gamma, const float* beta, float* y, float* stats, int32_t B, int32_t C, int32_t T,
                                void* stream) {
    TTTS_CHECK_ARG(x && gamma && beta && y && stats && B >= 1 && B <= 65535 && C >= 1 && T >= 1, "layernorm_c: bad args");
    TTTS_CUDA(launch_plain(lnc_stats_kernel, dim3((T + 31) / 32, B), dim3(256), 0, (cudaStream_t)stream, x, stats, C, T));
    TTTS_LAUNCH_CHECK("lnc_stats");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(lnc_apply_kernel, dim3(te_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, (const float*)stats, gamma, beta, y, C, T, n));
    TTTS_LAUNCH_CHECK("lnc_apply");
    return TTTS_OK;
}
/* dx [B,C,T], dgamma / dbeta [C] written ; scratch: B*T*2 floats ; stats from the forward */
extern "C" int ttts_layernorm_c_bwd(const float* dy, const float* x, const float* stats, const float* gamma, float* dx, float* dgamma, float* dbeta,
                                    float* scratch, int32_t B, int32_t C, int32_t T, void* stream) {
    TTTS_CHECK_ARG(dy && x && stats && gamma && dx && dgamma && dbeta && scratch && B >= 1 && B <= 65535 && C >= 1 && T >= 1, "layernorm_c backward: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CUDA(launch_plain(lnc_bwd_col_kernel, dim3((T + 31) / 32, B), dim3(256), 0, st, dy, x, stats, gamma, scratch, C, T));
    TTTS_LAUNCH_CHECK("lnc_bwd_col");
    const size_t n = (size_t)B * C * T;
    TTTS_CUDA(launch_plain(lnc_bwd_dx_kernel, dim3(te_blocks(n)), dim3(256), 0, st, dy, x, stats, (const float*)scratch, gamma, dx, C, T, n));
    TTTS_LAUNCH_CHECK("lnc_bwd_dx");
    TTTS_CUDA(launch_plain(lnc_bwd_param_kernel, dim3(C), dim3(256), 0, st, dy, x, stats, dgamma, dbeta, B, C, T));
    TTTS_LAUNCH_CHECK("lnc_bwd_param");
    return TTTS_OK;
}
