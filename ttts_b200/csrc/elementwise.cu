// Bandwidth-bound kernels of the GPT train step: token prep, embedding gather/scatter, LayerNorm fwd/bwd
// (single and the reference's double ln_f+final_norm), fused softmax-cross-entropy fwd/bwd, column sums
// (bias gradients), fp32->bf16 cast, global-norm clip + AdamW.  All are coalesced 128-bit sweeps, one warp
// per row, reduced with warp shuffles; they are HBM-bound by construction (DESIGN.md lists bytes/row).
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

// ------------------------------------------------------------------------------------------------
// token pre-processing: ttts/gpt/model.py:402-414 (set_mel_padding), 484-489 (start/stop padding)
// ------------------------------------------------------------------------------------------------
__global__ void prep_tokens_kernel(const int64_t* __restrict__ text, int ld_text, int64_t* __restrict__ codes, int ld_codes,
                                   const int64_t* __restrict__ wav_lengths, int B, int TL, int CL, int mel_comp,
                                   int start_text, int stop_text, int start_mel, int stop_mel,
                                   int32_t* __restrict__ text_in, int32_t* __restrict__ text_tgt,
                                   int32_t* __restrict__ mel_in, int32_t* __restrict__ mel_tgt) {
    const int b = blockIdx.x;
    const int Tt = TL + 2, Tm = CL + 2;
    // text: [start, t_0..t_{TL-1}, stop] ; target: [t_0.., stop, stop]
    for (int t = threadIdx.x; t < Tt; t += blockDim.x) {
        int vin = (t == 0) ? start_text : (t <= TL ? (int)text[(size_t)b * ld_text + t - 1] : stop_text);
        int vtg = (t < TL) ? (int)text[(size_t)b * ld_text + t] : stop_text;
        text_in[b * Tt + t] = vin;
        text_tgt[b * Tt + t] = vtg;
    }
    const int actual_end = (int)(wav_lengths[b] / mel_comp) + 1;
    for (int t = threadIdx.x; t < CL; t += blockDim.x) {
        if (t >= actual_end) codes[(size_t)b * ld_codes + t] = stop_mel;   // in-place on the caller's tensor (Appendix E #4)
    }
    __syncthreads();
    for (int t = threadIdx.x; t < Tm; t += blockDim.x) {
        int vin = (t == 0) ? start_mel : (t <= CL ? (int)codes[(size_t)b * ld_codes + t - 1] : stop_mel);
        int vtg = (t < CL) ? (int)codes[(size_t)b * ld_codes + t] : stop_mel;
        mel_in[b * Tm + t] = vin;
        mel_tgt[b * Tm + t] = vtg;
    }
}

int prep_tokens(const int64_t* text, int ld_text, int64_t* codes, int ld_codes, const int64_t* wav_lengths, int B, int TL, int CL,
                int mel_comp, int start_text, int stop_text, int start_mel, int stop_mel, int32_t* text_in, int32_t* text_tgt,
                int32_t* mel_in, int32_t* mel_tgt, cudaStream_t st) {
    prep_tokens_kernel<<<B, 256, 0, st>>>(text, ld_text, codes, ld_codes, wav_lengths, B, TL, CL, mel_comp, start_text, stop_text,
                                          start_mel, stop_mel, text_in, text_tgt, mel_in, mel_tgt);
    TTTS_LAUNCH_CHECK("prep_tokens");
    return TTTS_OK;
}

// ------------------------------------------------------------------------------------------------
// embeddings: x[b,t,:] = E[tok] + P[pos]  (+ embd dropout)      ttts/gpt/model.py:488,494-495,418 ; HF:modeling_gpt2.py drop
// ------------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const int32_t* __restrict__ text_in, const int32_t* __restrict__ mel_in,
                                 const float* __restrict__ Et, const float* __restrict__ Em, const float* __restrict__ Pt,
                                 const float* __restrict__ Pm, float* __restrict__ x, int B, int Tt, int Tm, int d, int Vt, int Vm,
                                 DropCfg drop) {
    const int T = Tt + Tm;
    const int row = blockIdx.x;
    const int b = row / T, t = row - b * T;
    const float *e, *p;
    if (t < Tt) {
        int tok = text_in[b * Tt + t];
        tok = min(max(tok, 0), Vt - 1);
        e = Et + (size_t)tok * d; p = Pt + (size_t)t * d;
    } else {
        int tok = mel_in[b * Tm + (t - Tt)];
        tok = min(max(tok, 0), Vm - 1);
        e = Em + (size_t)tok * d; p = Pm + (size_t)(t - Tt) * d;
    }
    float* o = x + (size_t)row * d;
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float4 a = *reinterpret_cast<const float4*>(e + c);
        float4 q = *reinterpret_cast<const float4*>(p + c);
        float v[4] = {a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w};
        if (drop.thresh16) {
            uint64_t bits = dropout_bits4(drop.seed, ((uint64_t)row * d + c) >> 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = dropout_keep(bits, j, drop.thresh16) ? v[j] * drop.scale : 0.f;
        }
        *reinterpret_cast<float4*>(o + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

int embed_fwd(const int32_t* text_in, const int32_t* mel_in, const float* Et, const float* Em, const float* Pt, const float* Pm,
              float* x, int B, int Tt, int Tm, int d, int Vt, int Vm, DropCfg drop, cudaStream_t st) {
    TTTS_CHECK_ARG(d % 4 == 0, "embed: d %% 4");
    int threads = d / 4 < 256 ? ((d / 4 + 31) / 32) * 32 : 256;
    embed_fwd_kernel<<<B * (Tt + Tm), threads, 0, st>>>(text_in, mel_in, Et, Em, Pt, Pm, x, B, Tt, Tm, d, Vt, Vm, drop);
    TTTS_LAUNCH_CHECK("embed_fwd");
    return TTTS_OK;
}

// token-table gradient: scatter-add rows of g (fp32 [B*T, d]) into dE ; position-table gradient: sum over batch.
__global__ void embed_bwd_tok_kernel(const int32_t* __restrict__ text_in, const int32_t* __restrict__ mel_in, const float* __restrict__ g,
                                     float* __restrict__ dEt, float* __restrict__ dEm, int B, int Tt, int Tm, int d, int Vt, int Vm,
                                     DropCfg drop) {
    const int T = Tt + Tm;
    const int row = blockIdx.x;
    const int b = row / T, t = row - b * T;
    float* dst;
    if (t < Tt) { int tok = min(max(text_in[b * Tt + t], 0), Vt - 1); dst = dEt + (size_t)tok * d; }
    else { int tok = min(max(mel_in[b * Tm + (t - Tt)], 0), Vm - 1); dst = dEm + (size_t)tok * d; }
    const float* src = g + (size_t)row * d;
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float4 a = *reinterpret_cast<const float4*>(src + c);
        float v[4] = {a.x, a.y, a.z, a.w};
        if (drop.thresh16) {
            uint64_t bits = dropout_bits4(drop.seed, ((uint64_t)row * d + c) >> 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = dropout_keep(bits, j, drop.thresh16) ? v[j] * drop.scale : 0.f;
        }
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
}
__global__ void embed_bwd_pos_kernel(const float* __restrict__ g, float* __restrict__ dPt, float* __restrict__ dPm, int B, int Tt, int Tm,
                                     int d, DropCfg drop) {
    const int T = Tt + Tm;
    const int t = blockIdx.x;
    float* dst = (t < Tt) ? dPt + (size_t)t * d : dPm + (size_t)(t - Tt) * d;
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int b = 0; b < B; ++b) {
            const size_t row = (size_t)b * T + t;
            float4 a = *reinterpret_cast<const float4*>(g + row * d + c);
            float v[4] = {a.x, a.y, a.z, a.w};
            if (drop.thresh16) {
                uint64_t bits = dropout_bits4(drop.seed, ((uint64_t)row * d + c) >> 2);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = dropout_keep(bits, j, drop.thresh16) ? v[j] * drop.scale : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += v[j];
        }
        float4 o = *reinterpret_cast<float4*>(dst + c);
        *reinterpret_cast<float4*>(dst + c) = make_float4(o.x + acc[0], o.y + acc[1], o.z + acc[2], o.w + acc[3]);
    }
}

int embed_bwd(const int32_t* text_in, const int32_t* mel_in, const float* g, float* dEt, float* dEm, float* dPt, float* dPm, int B, int Tt,
              int Tm, int d, int Vt, int Vm, DropCfg drop, cudaStream_t st) {
    int threads = d / 4 < 256 ? ((d / 4 + 31) / 32) * 32 : 256;
    embed_bwd_tok_kernel<<<B * (Tt + Tm), threads, 0, st>>>(text_in, mel_in, g, dEt, dEm, B, Tt, Tm, d, Vt, Vm, drop);
    TTTS_LAUNCH_CHECK("embed_bwd_tok");
    embed_bwd_pos_kernel<<<Tt + Tm, threads, 0, st>>>(g, dPt, dPm, B, Tt, Tm, d, drop);
    TTTS_LAUNCH_CHECK("embed_bwd_pos");
    return TTTS_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm.  One warp per row, the row lives in registers: lane owns float4 chunks at col = (i*32+lane)*4.
// NCH = d/128.  eps = 1e-5 (HF GPT2Config.layer_norm_epsilon, nn.LayerNorm default).
// ------------------------------------------------------------------------------------------------
constexpr float kLnEps = 1e-5f;

// [b, t] row order -> [all text rows ; all mel rows] (the two head GEMMs want contiguous row blocks)
TTTS_DEVICE int map_row(const RowMap& m, int row) {
    if (m.T == 0) return row;
    const int b = row / m.T, t = row - b * m.T;
    return t < m.Tt ? b * m.Tt + t : m.B * m.Tt + b * (m.T - m.Tt) + (t - m.Tt);
}

template <int NCH>
TTTS_DEVICE void ln_stats(const float (&v)[NCH * 4], int d, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * 4; ++i) s += v[i];
    mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH * 4; ++i) { float t = v[i] - mean; q += t * t; }
    rstd = rsqrtf(warp_sum(q) / (float)d + kLnEps);
}

// y = LN(x)*w+b ; DOUBLE: y = LN2(LN1(x))   (HF:modeling_gpt2.py:628 ln_f then ttts/gpt/model.py:427 final_norm)
template <int NCH, bool DOUBLE, bool OUT_BF16>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2, const float* __restrict__ b2, void* __restrict__ y,
                                                     float* __restrict__ stats, int M, RowMap map) {
    constexpr int d = NCH * 128;
    pdl_launch_dependents();                     // PDL (common.cuh)
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= M) return;
    const int orow = map_row(map, row);
    float v[NCH * 4];
    const float* xr = x + (size_t)row * d;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        float4 a = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
    float mean, rstd;
    ln_stats<NCH>(v, d, mean, rstd);
    if (lane == 0) { stats[(size_t)row * (DOUBLE ? 4 : 2)] = mean; stats[(size_t)row * (DOUBLE ? 4 : 2) + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 ww = __ldg(reinterpret_cast<const float4*>(w1 + c));
        float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + c));
        v[4 * i] = (v[4 * i] - mean) * rstd * ww.x + bb.x;
        v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * ww.y + bb.y;
        v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * ww.z + bb.z;
        v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * ww.w + bb.w;
    }
    if (DOUBLE) {
        ln_stats<NCH>(v, d, mean, rstd);
        if (lane == 0) { stats[(size_t)row * 4 + 2] = mean; stats[(size_t)row * 4 + 3] = rstd; }
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = (i * 32 + lane) * 4;
            float4 ww = __ldg(reinterpret_cast<const float4*>(w2 + c));
            float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + c));
            v[4 * i] = (v[4 * i] - mean) * rstd * ww.x + bb.x;
            v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * ww.y + bb.y;
            v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * ww.z + bb.z;
            v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * ww.w + bb.w;
        }
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (OUT_BF16) {
            uint2 o = make_uint2(pack_bf16(v[4 * i], v[4 * i + 1]), pack_bf16(v[4 * i + 2], v[4 * i + 3]));
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + (size_t)orow * d + c) = o;
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)orow * d + c) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    }
}

template <int NCH>
static int ln_fwd_launch(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, void* y, float* stats, int M,
                         bool dbl, bool out_bf16, RowMap map, cudaStream_t st) {
    dim3 grid((M + 7) / 8);
    if (dbl && out_bf16) TTTS_CUDA(launch_pdl(ln_fwd_kernel<NCH, true, true>, dim3(grid), dim3(256), 0, st, x, w1, b1, w2, b2, y, stats, M, map));
    else if (dbl) TTTS_CUDA(launch_pdl(ln_fwd_kernel<NCH, true, false>, dim3(grid), dim3(256), 0, st, x, w1, b1, w2, b2, y, stats, M, map));
    else if (out_bf16) TTTS_CUDA(launch_pdl(ln_fwd_kernel<NCH, false, true>, dim3(grid), dim3(256), 0, st, x, w1, b1, w2, b2, y, stats, M, map));
    else TTTS_CUDA(launch_pdl(ln_fwd_kernel<NCH, false, false>, dim3(grid), dim3(256), 0, st, x, w1, b1, w2, b2, y, stats, M, map));
    TTTS_LAUNCH_CHECK("ln_fwd");
    return TTTS_OK;
}

#define TTTS_LN_DISPATCH(FN, ...)                                                         \
    switch (d / 128) {                                                                   \
    case 1: return FN<1>(__VA_ARGS__);                                                   \
    case 2: return FN<2>(__VA_ARGS__);                                                   \
    case 3: return FN<3>(__VA_ARGS__);                                                   \
    case 4: return FN<4>(__VA_ARGS__);                                                   \
    case 6: return FN<6>(__VA_ARGS__);                                                   \
    case 8: return FN<8>(__VA_ARGS__);                                                   \
    default: ::ttts::set_error("layernorm: unsupported model_dim %d (need 128*{1,2,3,4,6,8})", d); return TTTS_ERR_INVALID; \
    }

int ln_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, void* y, float* stats, int M, int d,
           bool dbl, bool out_bf16, RowMap map, cudaStream_t st) {
    TTTS_CHECK_ARG(d % 128 == 0, "layernorm: d %% 128 != 0");
    TTTS_LN_DISPATCH(ln_fwd_launch, x, w1, b1, w2, b2, y, stats, M, dbl, out_bf16, map, st);
}

// Backward of y = LN(x)*w + b given dy, for one row held in registers.
//   dx = rstd * (dyw - mean(dyw) - xhat * mean(dyw*xhat)),  dyw = dy*w
// The per-row contributions to dgamma (dy*xhat) and dbeta (dy) are ACCUMULATED into this warp's own shared-memory rows (sdg, sdb):
// no block-level synchronisation in the row loop (the first version synchronised twice per 8 rows to reduce them, which kept the
// kernel at 58 % of HBM bandwidth), and the column accumulators do not live in registers (d/32 per array would be 96+).
template <int NCH>
TTTS_DEVICE void ln_bwd_row(const float (&xv)[NCH * 4], float mean, float rstd, const float* __restrict__ w, int lane, int d,
                            float (&dy)[NCH * 4] /* in: dy ; out: dx */, float* __restrict__ sdg, float* __restrict__ sdb) {
    float s1 = 0.f, s2 = 0.f;
    float xh[NCH * 4];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
        float wv[4] = {ww.x, ww.y, ww.z, ww.w};
        float pg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * i + j;
            xh[k] = (xv[k] - mean) * rstd;
            pg[j] = dy[k] * xh[k];
        }
        float4 ag = *reinterpret_cast<const float4*>(sdg + c), ab = *reinterpret_cast<const float4*>(sdb + c);
        ag.x += pg[0]; ag.y += pg[1]; ag.z += pg[2]; ag.w += pg[3];
        ab.x += dy[4 * i]; ab.y += dy[4 * i + 1]; ab.z += dy[4 * i + 2]; ab.w += dy[4 * i + 3];
        *reinterpret_cast<float4*>(sdg + c) = ag;
        *reinterpret_cast<float4*>(sdb + c) = ab;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * i + j;
            dy[k] *= wv[j];
            s1 += dy[k];
            s2 += dy[k] * xh[k];
        }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
#pragma unroll
    for (int k = 0; k < NCH * 4; ++k) dy[k] = rstd * (dy[k] - s1 - xh[k] * s2);
}

// g_out = g_in + LNbwd(dy) (fp32, may alias g_in; g_in may be NULL) ; g16_out = bf16(dropmask(g_out)) for the next
// dgrad GEMM ; column sums: dgamma/dbeta of this LN, and dbias_next = colsum(g16_out) (bias of the projection whose
// output was added to the residual just below this LN).  DOUBLE: through final_norm then ln_f.
// dynamic smem: NARR * 8 * d floats, NARR = 3 (single) or 5 (double).
template <int NCH, bool DOUBLE>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const void* __restrict__ dy_in, int dy_is_f32, const float* __restrict__ x,
                                                     const float* __restrict__ stats, const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2, const float* g_in, float* g_out, bf16* __restrict__ g16_out,
                                                     float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                                     float* __restrict__ db2, float* __restrict__ dbias_next, int M, DropCfg drop, RowMap map) {
    constexpr int d = NCH * 128;
    constexpr int NARR = DOUBLE ? 5 : 3;
    extern __shared__ float ln_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_row[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) s_row[a] = ln_smem + ((size_t)a * 8 + warp) * d;
    // arrays: 0 = dgamma1, 1 = dbeta1, 2 = dbias_next, 3 = dgamma2, 4 = dbeta2
    float acc[NARR][4];
#pragma unroll
    for (int a = 0; a < NARR; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;

#pragma unroll
    for (int a = 0; a < NARR; ++a)
        for (int c = lane * 4; c < d; c += 128) *reinterpret_cast<float4*>(s_row[a] + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    pdl_launch_dependents();                     // PDL (common.cuh): the shared-memory clearing above overlapped the previous kernel's tail
    pdl_wait();
    for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
        {
            float xv[NCH * 4], dy[NCH * 4];
            const float* xr = x + (size_t)row * d;
            const int drow = map_row(map, row);
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int c = (i * 32 + lane) * 4;
                float4 a = *reinterpret_cast<const float4*>(xr + c);
                xv[4 * i] = a.x; xv[4 * i + 1] = a.y; xv[4 * i + 2] = a.z; xv[4 * i + 3] = a.w;
                if (dy_is_f32) {
                    float4 g = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_in) + (size_t)drow * d + c);
                    dy[4 * i] = g.x; dy[4 * i + 1] = g.y; dy[4 * i + 2] = g.z; dy[4 * i + 3] = g.w;
                } else {
                    uint2 g = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy_in) + (size_t)drow * d + c);
                    dy[4 * i] = bf16_lo(g.x); dy[4 * i + 1] = bf16_hi(g.x); dy[4 * i + 2] = bf16_lo(g.y); dy[4 * i + 3] = bf16_hi(g.y);
                }
            }
            if (DOUBLE) {
                const float mean1 = stats[(size_t)row * 4], rstd1 = stats[(size_t)row * 4 + 1];
                const float mean2 = stats[(size_t)row * 4 + 2], rstd2 = stats[(size_t)row * 4 + 3];
                float z[NCH * 4];
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const int c = (i * 32 + lane) * 4;
                    float4 ww = __ldg(reinterpret_cast<const float4*>(w1 + c));
                    float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + c));
                    z[4 * i] = (xv[4 * i] - mean1) * rstd1 * ww.x + bb.x;
                    z[4 * i + 1] = (xv[4 * i + 1] - mean1) * rstd1 * ww.y + bb.y;
                    z[4 * i + 2] = (xv[4 * i + 2] - mean1) * rstd1 * ww.z + bb.z;
                    z[4 * i + 3] = (xv[4 * i + 3] - mean1) * rstd1 * ww.w + bb.w;
                }
                ln_bwd_row<NCH>(z, mean2, rstd2, w2, lane, d, dy, s_row[DOUBLE ? 3 : 0], s_row[DOUBLE ? 4 : 1]);   // dy := dz
                ln_bwd_row<NCH>(xv, mean1, rstd1, w1, lane, d, dy, s_row[0], s_row[1]);                            // dy := dx
            } else {
                const float mean = stats[(size_t)row * 2], rstd = stats[(size_t)row * 2 + 1];
                ln_bwd_row<NCH>(xv, mean, rstd, w1, lane, d, dy, s_row[0], s_row[1]);
            }
            // g_in may alias g_out: with the load inside the store loop the compiler must keep load(i + 1) behind store(i) -- NCH dependent
            // round trips to HBM per row.  All loads of the row are issued before its first store instead (xv is dead, its registers are reused).
            if (g_in) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const float4 gi = *reinterpret_cast<const float4*>(g_in + (size_t)row * d + (i * 32 + lane) * 4);
                    dy[4 * i] += gi.x; dy[4 * i + 1] += gi.y; dy[4 * i + 2] += gi.z; dy[4 * i + 3] += gi.w;
                }
            }
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int c = (i * 32 + lane) * 4;
                float o[4] = {dy[4 * i], dy[4 * i + 1], dy[4 * i + 2], dy[4 * i + 3]};
                *reinterpret_cast<float4*>(g_out + (size_t)row * d + c) = make_float4(o[0], o[1], o[2], o[3]);
                if (g16_out) {
                    if (drop.thresh16) {
                        uint64_t bits = dropout_bits4(drop.seed, ((uint64_t)row * d + c) >> 2);
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = dropout_keep(bits, j, drop.thresh16) ? o[j] * drop.scale : 0.f;
                    }
                    uint2 pk = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
                    *reinterpret_cast<uint2*>(g16_out + (size_t)row * d + c) = pk;
                    float4 bn = *reinterpret_cast<const float4*>(s_row[2] + c);
                    bn.x += bf16_lo(pk.x); bn.y += bf16_hi(pk.x); bn.z += bf16_lo(pk.y); bn.w += bf16_hi(pk.y);
                    *reinterpret_cast<float4*>(s_row[2] + c) = bn;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x * 4 < d) {
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float4 t = *reinterpret_cast<const float4*>(ln_smem + ((size_t)a * 8 + r) * d + threadIdx.x * 4);
                acc[a][0] += t.x; acc[a][1] += t.y; acc[a][2] += t.z; acc[a][3] += t.w;
            }
        }
    }
    if (threadIdx.x * 4 < d) {
        float* dst[5] = {dw1, db1, dbias_next, dw2, db2};
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            if (dst[a] == nullptr) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dst[a] + threadIdx.x * 4 + j, acc[a][j]);
        }
    }
}

template <int NCH>
static int ln_bwd_launch(const void* dy, int dy_is_f32, const float* x, const float* stats, const float* w1, const float* b1, const float* w2,
                         const float* g_in, float* g_out, bf16* g16_out, float* dw1, float* db1, float* dw2, float* db2, float* dbias_next,
                         int M, bool dbl, DropCfg drop, RowMap map, cudaStream_t st) {
    int blocks = num_sms() * 2;
    int need = (M + 7) / 8;
    if (blocks > need) blocks = need;
    const int d = NCH * 128;
    const size_t smem = (size_t)(dbl ? 5 : 3) * 8 * d * sizeof(float);
    if (dbl) {
        static bool attr = false;
        if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 8 * d * 4)); attr = true; }
        TTTS_CUDA(launch_pdl(ln_bwd_kernel<NCH, true>, dim3(blocks), dim3(256), smem, st, dy, dy_is_f32, x, stats, w1, b1, w2, g_in, g_out, g16_out, dw1, db1, dw2, db2, dbias_next, M, drop, map));
    } else {
        static bool attr = false;
        if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<NCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 8 * d * 4)); attr = true; }
        TTTS_CUDA(launch_pdl(ln_bwd_kernel<NCH, false>, dim3(blocks), dim3(256), smem, st, dy, dy_is_f32, x, stats, w1, b1, w2, g_in, g_out, g16_out, dw1, db1, dw2, db2, dbias_next, M, drop, map));
    }
    TTTS_LAUNCH_CHECK("ln_bwd");
    return TTTS_OK;
}

int ln_bwd(const void* dy, int dy_is_f32, const float* x, const float* stats, const float* w1, const float* b1, const float* w2,
           const float* g_in, float* g_out, bf16* g16_out, float* dw1, float* db1, float* dw2, float* db2, float* dbias_next, int M, int d,
           bool dbl, DropCfg drop, RowMap map, cudaStream_t st) {
    TTTS_CHECK_ARG(d % 128 == 0, "layernorm: d %% 128 != 0");
    TTTS_LN_DISPATCH(ln_bwd_launch, dy, dy_is_f32, x, stats, w1, b1, w2, g_in, g_out, g16_out, dw1, db1, dw2, db2, dbias_next, M, dbl, drop, map, st);
}

// ------------------------------------------------------------------------------------------------
// fused softmax cross-entropy over bf16 logits [rows, ld] (V valid columns)   ttts/gpt/model.py:508-509
//   fwd: row_loss[r] = lse - logit[target], lse saved ; bwd: dlogits = (softmax - onehot) * (*gscale) / rows_total
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_fwd_kernel(const bf16* __restrict__ logits, int ld, int V, const int32_t* __restrict__ tgt, int rows,
                                                     float* __restrict__ row_loss, float* __restrict__ row_lse) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const bf16* lr = logits + (size_t)row * ld;
    float m = -INFINITY;
    for (int c = lane * 2; c < V; c += 64) {
        uint32_t u = *reinterpret_cast<const uint32_t*>(lr + c);
        m = fmaxf(m, bf16_lo(u));
        if (c + 1 < V) m = fmaxf(m, bf16_hi(u));
    }
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane * 2; c < V; c += 64) {
        uint32_t u = *reinterpret_cast<const uint32_t*>(lr + c);
        s += expf(bf16_lo(u) - m);
        if (c + 1 < V) s += expf(bf16_hi(u) - m);
    }
    s = warp_sum(s);
    if (lane == 0) {
        const float lse = m + logf(s);
        int t = min(max(tgt[row], 0), V - 1);
        row_lse[row] = lse;
        row_loss[row] = lse - __bfloat162float(lr[t]);
    }
}
// deterministic mean of n values -> out[0]
__global__ void __launch_bounds__(1024) mean_reduce_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float sm[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = s / (float)n;
    }
}
__global__ void __launch_bounds__(256) ce_bwd_kernel(const bf16* __restrict__ logits, int ld, int V, const int32_t* __restrict__ tgt, int rows,
                                                     const float* __restrict__ row_lse, const float* __restrict__ gscale, float weight,
                                                     bf16* __restrict__ dlogits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const bf16* lr = logits + (size_t)row * ld;
    bf16* dr = dlogits + (size_t)row * ld;
    const float lse = row_lse[row];
    const float g = (gscale ? gscale[0] : 1.0f) * weight / (float)rows;
    const int t = min(max(tgt[row], 0), V - 1);
    for (int c = lane * 2; c < ld; c += 64) {
        uint32_t u = *reinterpret_cast<const uint32_t*>(lr + c);
        float p0 = (c < V) ? (expf(bf16_lo(u) - lse) - (c == t ? 1.f : 0.f)) * g : 0.f;
        float p1 = (c + 1 < V) ? (expf(bf16_hi(u) - lse) - (c + 1 == t ? 1.f : 0.f)) * g : 0.f;
        *reinterpret_cast<uint32_t*>(dr + c) = pack_bf16(p0, p1);
    }
}

int ce_fwd(const bf16* logits, int ld, int V, const int32_t* tgt, int rows, float* row_loss, float* row_lse, float* loss_out, cudaStream_t st) {
    TTTS_CHECK_ARG(ld % 2 == 0, "ce: ld must be even");
    ce_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(logits, ld, V, tgt, rows, row_loss, row_lse);
    TTTS_LAUNCH_CHECK("ce_fwd");
    mean_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, rows, loss_out);
    TTTS_LAUNCH_CHECK("mean_reduce");
    return TTTS_OK;
}
int ce_bwd(const bf16* logits, int ld, int V, const int32_t* tgt, int rows, const float* row_lse, const float* gscale, float weight,
           bf16* dlogits, cudaStream_t st) {
    TTTS_CHECK_ARG(ld % 2 == 0, "ce: ld must be even");
    ce_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(logits, ld, V, tgt, rows, row_lse, gscale, weight, dlogits);
    TTTS_LAUNCH_CHECK("ce_bwd");
    return TTTS_OK;
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix [M, ld] (N columns) -> out[N] += sum_rows   (bias gradients)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ a, int ld, int M, int N, float* __restrict__ out) {
    pdl_launch_dependents();                     // PDL (common.cuh)
    pdl_wait();
    // block = 32 column groups of 8 (one 16-byte load each) x 8 row lanes ; grid.x over column tiles of 256, grid.y over row chunks.
    // Four independent 16-byte loads per thread in flight (the first version moved 4 bytes per load with one load in flight).
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + cx * 8;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    auto add8 = [&](const uint4 u) {
        s[0] += bf16_lo(u.x); s[1] += bf16_hi(u.x); s[2] += bf16_lo(u.y); s[3] += bf16_hi(u.y);
        s[4] += bf16_lo(u.z); s[5] += bf16_hi(u.z); s[6] += bf16_lo(u.w); s[7] += bf16_hi(u.w);
    };
    if (col + 8 <= N) {
        int r = r0 + ry;
        for (; r + 24 < r1; r += 32) {
            const uint4 u0 = *reinterpret_cast<const uint4*>(a + (size_t)r * ld + col);
            const uint4 u1 = *reinterpret_cast<const uint4*>(a + (size_t)(r + 8) * ld + col);
            const uint4 u2 = *reinterpret_cast<const uint4*>(a + (size_t)(r + 16) * ld + col);
            const uint4 u3 = *reinterpret_cast<const uint4*>(a + (size_t)(r + 24) * ld + col);
            add8(u0); add8(u1); add8(u2); add8(u3);
        }
        for (; r < r1; r += 8) add8(*reinterpret_cast<const uint4*>(a + (size_t)r * ld + col));
    } else if (col < N) {                                   // ragged last column group
        for (int r = r0 + ry; r < r1; r += 8)
            for (int k = 0; k < 8 && col + k < N; ++k) s[k] += __bfloat162float(a[(size_t)r * ld + col + k]);
    }
    __shared__ float sm[8][256 + 8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[ry][cx * 8 + k] = s[k];
    __syncthreads();
    {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
        const int c = blockIdx.x * 256 + threadIdx.x;
        if (c < N) atomicAdd(out + c, t);
    }
}
int colsum_bf16(const bf16* a, int ld, int M, int N, float* out, cudaStream_t st) {
    TTTS_CHECK_ARG(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0, "colsum: ld must be a multiple of 8 and the matrix 16-byte aligned");
    int gx = (N + 255) / 256;
    int gy = (num_sms() * 8 + gx - 1) / gx;
    if (gy > (M + 63) / 64) gy = (M + 63) / 64;
    if (gy < 1) gy = 1;
    TTTS_CUDA(launch_pdl(colsum_bf16_kernel, dim3(gx, gy), dim3(256), 0, st, a, ld, M, N, out));
    TTTS_LAUNCH_CHECK("colsum_bf16");
    return TTTS_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast of the flat parameter buffer
// ------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 a = reinterpret_cast<const float4*>(src)[i];
        reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
    }
}
int cast_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st) {
    TTTS_CHECK_ARG(n % 4 == 0, "cast: n %% 4 != 0");
    size_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cast_bf16_kernel<<<blocks, 256, 0, st>>>(src, dst, n4);
    TTTS_LAUNCH_CHECK("cast_bf16");
    return TTTS_OK;
}

// ------------------------------------------------------------------------------------------------
// global-norm clip + AdamW over the flat buffers   (ttts/gpt/train.py:22-31,114-118 ; torch.optim.AdamW)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, size_t n4, float* __restrict__ partial) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float s = 0.f;
    for (; i < n4; i += stride) {
        float4 a = reinterpret_cast<const float4*>(g)[i];
        s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    __shared__ float sm[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(1024) sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out_norm) {
    __shared__ double sm[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = sm[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out_norm[0] = (float)sqrt(s);
    }
}
int grad_norm(const float* g, size_t n, float* partial /*[>=1024]*/, float* out_norm, cudaStream_t st) {
    TTTS_CHECK_ARG(n % 4 == 0, "grad_norm: n %% 4 != 0");
    int blocks = num_sms() * 4;
    if (blocks > 1024) blocks = 1024;
    sumsq_partial_kernel<<<blocks, 256, 0, st>>>(g, n / 4, partial);
    TTTS_LAUNCH_CHECK("sumsq_partial");
    sumsq_final_kernel<<<1, 1024, 0, st>>>(partial, blocks, out_norm);
    TTTS_LAUNCH_CHECK("sumsq_final");
    return TTTS_OK;
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    bf16* __restrict__ p16, size_t n4, const float* __restrict__ norm, float max_norm,
                                                    float grad_scale, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                                                    float bc2_sqrt) {
    // torch clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max=1)
    float coef = grad_scale;
    if (norm != nullptr && max_norm > 0.f) coef *= fminf(1.0f, max_norm / (norm[0] * grad_scale + 1e-6f));
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float P[4] = {pp.x, pp.y, pp.z, pp.w}, G[4] = {gg.x, gg.y, gg.z, gg.w}, Mv[4] = {mm.x, mm.y, mm.z, mm.w}, V[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = G[j] * coef;
            P[j] *= (1.0f - lr * wd);
            Mv[j] = beta1 * Mv[j] + (1.0f - beta1) * gj;
            V[j] = beta2 * V[j] + (1.0f - beta2) * gj * gj;
            const float denom = sqrtf(V[j]) / bc2_sqrt + eps;
            P[j] -= (lr / bc1) * (Mv[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(P[0], P[1], P[2], P[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(Mv[0], Mv[1], Mv[2], Mv[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(V[0], V[1], V[2], V[3]);
        if (p16) reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_bf16(P[0], P[1]), pack_bf16(P[2], P[3]));
    }
}
int adamw_step(float* p, const float* g, float* m, float* v, bf16* p16, size_t n, const float* norm, float max_norm, float grad_scale,
               float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t st) {
    TTTS_CHECK_ARG(n % 4 == 0 && step >= 1, "adamw: bad n/step");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    size_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    int cap = num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adamw_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, p16, n4, norm, max_norm, grad_scale, lr, beta1, beta2, eps, wd, bc1, sqrtf(bc2));
    TTTS_LAUNCH_CHECK("adamw");
    return TTTS_OK;
}

}  // namespace ttts
