// KV-cache decode step of the UnifiedVoice GPT: one new code per sequence against cached keys / values -- the cached branch of the
// reference's GPT2InferenceModel (ttts/gpt/model.py:34-200: `past_key_values`, one input id per step, :144-147) as it is driven by
// `inference_speech` (:533-562) through HF generate.  SURVEY.md 8(f) #4, last item.
//
// What bounds it: with B = 1..8 sequences every weight byte is used B times, so a step is a sweep over the bf16 parameter shadow
// (24 d^2 bytes per layer: 604 MB at L24 / d1024 -> 78 us at the measured HBM peak) plus the cache rows of each (sequence, head)
// (256 B per cached position per layer).  No tensor cores: the "GEMMs" are M <= 8 matrix-vector sweeps, HBM-bound by construction.
//
// Arithmetic mirrors the train-step kernels operand for operand, so cached and uncached generation agree up to fp32 summation order:
// LayerNorm in fp32 on the fp32 residual stream -> bf16 ; bf16 x bf16 products accumulated in fp32 ; bf16(acc + bias) where the GEMM
// epilogues round (c_attn out, both c_proj outs before the residual add, c_fc pre-activation) ; gelu_new in packed bf16x2 ; attention
// probabilities rounded to bf16 before P.V with an fp32 normaliser ; heads: bf16(acc + bias).
//
// Kernels per layer (8 launches; all small, all with programmatic dependent launch so the launch latencies overlap):
//   dec_rowop   (finalise the previous projection into the residual stream, or embed; LayerNorm -> bf16)      grid B
//   dec_gemv    c_attn   partial sums over a K slice: part[s][b][n]                                           grid (N/256, S, B/BT)
//   dec_attn    sum partials + bias -> q,k,v ; append k,v to the cache ; softmax(q K^T / 8) V                 grid (H, B)
//   dec_gemv    attn c_proj
//   dec_rowop   residual add + ln_2
//   dec_gemv    c_fc ; dec_act (sum partials + bias -> bf16 -> gelu_new)
//   dec_gemv    mlp c_proj
// then dec_rowop (residual add + ln_f + final_norm), dec_head (mel head), dec_advance (slot += 1).
// Every kernel reads the current cache slot from DEVICE memory, so one recorded step (a CUDA graph) replays for every position.
//
// This file is also compiled for the HOST by tests/emu (g++ -DTTTS_HOST_EMU: one OS thread per CUDA thread, blocks one after another),
// which runs these very kernels and the launch sequence below on CPU against the oracle -- hence no <<<>>> and the TTTS_DYN_SMEM macro.
#include <string.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#define TTTS_DYN_SMEM(type, name) extern __shared__ type name[]
#endif
#include "gpt_layout.h"

namespace ttts {

constexpr float kDecLnEps = 1e-5f;
constexpr int DEC_SPLIT_MAX = 16;      // K slices of a matrix-vector sweep (partials are summed by the consumer in slice order: deterministic)
constexpr int DEC_PF = 8;              // weight rows per warp requested ahead of griddepcontrol.wait in the sweep
constexpr int DEC_BT = 4;              // sequences per CTA of the sweep (weights are re-read from L2 for the next group)

// block-wide sum for 256 threads (8 warps); `red` holds >= 8 floats; safe to call back to back (trailing barrier)
TTTS_DEVICE float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    __syncthreads();
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// dec_rowop: x = embed(token, position)                      (mode 0)       ttts/gpt/model.py:134-147
//            x = resid + bf16(bias + sum_s part[s])          (mode 1)       HF: modeling_gpt2.py:282,307 (dropout off)
//        then resid = x ; xn16 = bf16(LN(x))  or  bf16(LN2(LN1(x)))         HF: :273,304,628 ; ttts/gpt/model.py:427
// One CTA per sequence, 256 threads, d <= 1024 -> <= 4 columns per thread (c = tid + 256 i).
// ---------------------------------------------------------------------------------------------------------------------------------
struct RowopArgs {
    int mode, d, B, S;
    const float* part;            // [S][B][d]
    const float* bias;            // [d]
    float* resid;                 // [B][d]
    const float *w1, *b1, *w2, *b2;   // w2 == nullptr: single LayerNorm
    bf16* xn16;                   // [B][d]
    // mode 0
    const int64_t* codes; int ld_codes;
    const int32_t* slot; int text_positions, pos_shift, T_max, Vm, n_pos_rows;
    const float *Em, *Pm;
};

__global__ void __launch_bounds__(256) dec_rowop_kernel(const RowopArgs a) {
    __shared__ float red[8];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x, d = a.d;
    float v[4];
    if (a.mode == 0) {
        const int slot = *a.slot;
        if (slot >= a.T_max || slot <= a.text_positions) return;     // host checks capacity; never write out of bounds
        const int j = slot - a.text_positions;                        // index inside the mel segment (start_mel is 0)
        int tok = (int)a.codes[(size_t)b * a.ld_codes + (j - 1)];
        tok = min(max(tok, 0), a.Vm - 1);
        const int pos = min(j + a.pos_shift, a.n_pos_rows - 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + 256 * i;
            v[i] = c < d ? a.Em[(size_t)tok * d + c] + a.Pm[(size_t)pos * d + c] : 0.f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + 256 * i;
            float acc = 0.f;
            if (c < d) {
                // all slices requested before the first add (a run-time loop made 16 DEPENDENT L2 round trips: dec_rowop was 20 us, r2s)
                float ps[DEC_SPLIT_MAX];
#pragma unroll
                for (int s = 0; s < DEC_SPLIT_MAX; ++s) ps[s] = s < a.S ? a.part[((size_t)s * a.B + b) * d + c] : 0.f;
                const float rv = a.resid[(size_t)b * d + c], bv = a.bias[c];
#pragma unroll
                for (int s = 0; s < DEC_SPLIT_MAX; ++s) acc += ps[s];
                acc = rv + bf16_round(acc + bv);
            }
            v[i] = acc;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = tid + 256 * i;
        if (c < d) a.resid[(size_t)b * d + c] = v[i];
    }
    for (int pass = 0; pass < 2; ++pass) {
        const float* w = pass == 0 ? a.w1 : a.w2;
        const float* bb = pass == 0 ? a.b1 : a.b2;
        if (w == nullptr) break;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += (tid + 256 * i < d) ? v[i] : 0.f;
        const float mean = block_sum_256(s, red) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float t = v[i] - mean; q += (tid + 256 * i < d) ? t * t : 0.f; }
        const float rstd = rsqrtf(block_sum_256(q, red) / (float)d + kDecLnEps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + 256 * i;
            if (c < d) v[i] = (v[i] - mean) * rstd * w[c] + bb[c];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = tid + 256 * i;
        if (c < d) a.xn16[(size_t)b * d + c] = __float2bfloat16_rn(v[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// dec_gemv: part[s][b][n] = sum_{k in slice s} x16[b][k] * W16[k][n]      W = HF Conv1D weight [K][N] (N contiguous), HF: pytorch_utils.py:119-123
// CTA = 256 columns x one K slice x BT sequences.  Warp w owns rows k = w, w + 8, ... of the slice; lane owns 8 consecutive columns
// (one 16-byte load, a warp reads 512 contiguous bytes of a weight row); x[b][k] is a shared-memory broadcast.  The 8 per-warp partial
// rows are summed through shared memory in warp order, so the result does not depend on scheduling.
// ---------------------------------------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(256) dec_gemv_kernel(const bf16* __restrict__ x16, const bf16* __restrict__ W, float* __restrict__ part, int K, int N,
                                                       int B, int ks) {
    TTTS_DYN_SMEM(float, dec_smem);
    float* xs = dec_smem;                       // [BT][ks]
    float* red = dec_smem + BT * ks;            // [8][BT][256]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = blockIdx.y * ks, b0 = blockIdx.z * BT;
    const int n0 = blockIdx.x * 256 + lane * 8;
    // The weights do not depend on the previous kernel: the first DEC_PF rows of this warp are requested BEFORE griddepcontrol.wait, so the
    // weight stream -- the only HBM traffic of a decode step that matters -- overlaps the tail of the producer instead of starting after it
    // (r2s launch list: 8.3 us per sweep for 2 - 8 MB of weights, i.e. latency, not bandwidth).
    const bf16* wp = W + (size_t)k0 * N + n0;
    uint4 wpre[DEC_PF];
#pragma unroll
    for (int j = 0; j < DEC_PF; ++j) {
        const int k = warp + 8 * j;
        wpre[j] = (n0 < N && k < ks) ? __ldg(reinterpret_cast<const uint4*>(wp + (size_t)k * N)) : make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_launch_dependents();
    pdl_wait();
    for (int i = tid; i < BT * ks; i += 256) {
        const int bt = i / ks, k = i - bt * ks;
        xs[i] = (b0 + bt < B) ? __bfloat162float(x16[(size_t)(b0 + bt) * K + k0 + k]) : 0.f;
    }
    __syncthreads();
    float acc[BT][8];
#pragma unroll
    for (int bt = 0; bt < BT; ++bt)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[bt][j] = 0.f;
    if (n0 < N) {
#pragma unroll
        for (int jp = 0; jp < DEC_PF; ++jp) {
            const int k = warp + 8 * jp;
            if (k < ks) {
                const uint4 w = wpre[jp];
                const float wf[8] = {bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y), bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w)};
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) {
                    const float xv = xs[bt * ks + k];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[bt][j] = fmaf(xv, wf[j], acc[bt][j]);
                }
            }
        }
#pragma unroll 4
        for (int k = warp + 8 * DEC_PF; k < ks; k += 8) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)k * N));
            const float wf[8] = {bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y), bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w)};
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float xv = xs[bt * ks + k];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[bt][j] = fmaf(xv, wf[j], acc[bt][j]);
            }
        }
    }
#pragma unroll
    for (int bt = 0; bt < BT; ++bt)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(warp * BT + bt) * 256 + lane * 8 + j] = acc[bt][j];
    __syncthreads();
    const int n = blockIdx.x * 256 + tid;
    if (n < N) {
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) {
            if (b0 + bt >= B) break;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[(w * BT + bt) * 256 + tid];
            part[((size_t)blockIdx.y * B + b0 + bt) * N + n] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// dec_gemv_fused: the sweep with its INPUT computed in the prologue of every CTA instead of by a kernel of its own (r2t launch list: after the
// latency fixes a step is still ~200 launches of ~5 us; dec_rowop and dec_act are 2 + 1 of the 8 launches of a layer and do microseconds of work).
//   MODE 1: x = bf16(LN(resid + bf16(bias + sum_s part_in[s]))) -- the residual add + LayerNorm of dec_rowop.  Every CTA needs the whole row for
//           the statistics (64 KB of L2-resident partials); CTA (0, 0, z) also writes the new residual row, into the OTHER residual buffer
//           (the CTAs of one launch read the old one at different times).
//   MODE 2: x = gelu_new(bf16(bias + sum_s part_in[s])) of this CTA's K slice -- dec_act.
// Same arithmetic and rounding points as the separate kernels: bit-identical results.  part_in and part are different buffers.
// ---------------------------------------------------------------------------------------------------------------------------------
struct GemvPro {
    const float* part_in; int S_in;      // [S_in][B][K]
    const float* bias_in;               // [K]
    const float* resid_in; float* resid_out; const float *ln_w, *ln_b;      // MODE 1
};

template <int BT, int MODE>
__global__ void __launch_bounds__(256) dec_gemv_fused_kernel(const GemvPro g, const bf16* __restrict__ W, float* __restrict__ part, int K, int N, int B, int ks) {
    TTTS_DYN_SMEM(float, dec_smem);
    __shared__ float red8[8];
    float* xs = dec_smem;                       // [BT][ks]
    float* red = dec_smem + BT * ks;            // [8][BT][256]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = blockIdx.y * ks, b0 = blockIdx.z * BT;
    const int n0 = blockIdx.x * 256 + lane * 8;
    const bf16* wp = W + (size_t)k0 * N + n0;
    uint4 wpre[DEC_PF];
#pragma unroll
    for (int j = 0; j < DEC_PF; ++j) {
        const int k = warp + 8 * j;
        wpre[j] = (n0 < N && k < ks) ? __ldg(reinterpret_cast<const uint4*>(wp + (size_t)k * N)) : make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_launch_dependents();
    pdl_wait();
    if (MODE == 1) {
        const bool writer = blockIdx.x == 0 && blockIdx.y == 0;
        for (int bt = 0; bt < BT; ++bt) {
            const int b = b0 + bt;
            if (b >= B) {                                             // uniform per CTA
                for (int k = tid; k < ks; k += 256) xs[bt * ks + k] = 0.f;
                continue;
            }
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid + 256 * i;
                float acc = 0.f;
                if (c < K) {
                    float ps[DEC_SPLIT_MAX];
#pragma unroll
                    for (int sl = 0; sl < DEC_SPLIT_MAX; ++sl) ps[sl] = sl < g.S_in ? g.part_in[((size_t)sl * B + b) * K + c] : 0.f;
                    const float rv = g.resid_in[(size_t)b * K + c], bv = g.bias_in[c];
#pragma unroll
                    for (int sl = 0; sl < DEC_SPLIT_MAX; ++sl) acc += ps[sl];
                    acc = rv + bf16_round(acc + bv);
                    if (writer) g.resid_out[(size_t)b * K + c] = acc;
                }
                v[i] = acc;
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) sum += (tid + 256 * i < K) ? v[i] : 0.f;
            const float mean = block_sum_256(sum, red8) / (float)K;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float t = v[i] - mean; q += (tid + 256 * i < K) ? t * t : 0.f; }
            const float rstd = rsqrtf(block_sum_256(q, red8) / (float)K + kDecLnEps);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = tid + 256 * i;
                if (c >= k0 && c < k0 + ks) xs[bt * ks + c - k0] = bf16_round((v[i] - mean) * rstd * g.ln_w[c] + g.ln_b[c]);
            }
        }
    } else {
        const int half = ks >> 1;
        for (int i = tid; i < BT * half; i += 256) {
            const int bt = i / half, kp = i - bt * half, b = b0 + bt;
            float x0 = 0.f, x1 = 0.f;
            if (b < B) {
                const int n = k0 + 2 * kp;
                float2 ps[DEC_SPLIT_MAX];
#pragma unroll
                for (int sl = 0; sl < DEC_SPLIT_MAX; ++sl)
                    ps[sl] = sl < g.S_in ? *reinterpret_cast<const float2*>(g.part_in + ((size_t)sl * B + b) * K + n) : make_float2(0.f, 0.f);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int sl = 0; sl < DEC_SPLIT_MAX; ++sl) { a0 += ps[sl].x; a1 += ps[sl].y; }
                const uint32_t r = gelu_new_bf2(pack_bf16(a0 + g.bias_in[n], a1 + g.bias_in[n + 1]));
                x0 = bf16_lo(r); x1 = bf16_hi(r);
            }
            xs[bt * ks + 2 * kp] = x0; xs[bt * ks + 2 * kp + 1] = x1;
        }
    }
    __syncthreads();
    float acc[BT][8];
#pragma unroll
    for (int bt = 0; bt < BT; ++bt)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[bt][j] = 0.f;
    if (n0 < N) {
#pragma unroll
        for (int jp = 0; jp < DEC_PF; ++jp) {
            const int k = warp + 8 * jp;
            if (k < ks) {
                const uint4 w = wpre[jp];
                const float wf[8] = {bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y), bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w)};
#pragma unroll
                for (int bt = 0; bt < BT; ++bt) {
                    const float xv = xs[bt * ks + k];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[bt][j] = fmaf(xv, wf[j], acc[bt][j]);
                }
            }
        }
#pragma unroll 4
        for (int k = warp + 8 * DEC_PF; k < ks; k += 8) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)k * N));
            const float wf[8] = {bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y), bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w)};
#pragma unroll
            for (int bt = 0; bt < BT; ++bt) {
                const float xv = xs[bt * ks + k];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[bt][j] = fmaf(xv, wf[j], acc[bt][j]);
            }
        }
    }
#pragma unroll
    for (int bt = 0; bt < BT; ++bt)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(warp * BT + bt) * 256 + lane * 8 + j] = acc[bt][j];
    __syncthreads();
    const int n = blockIdx.x * 256 + tid;
    if (n < N) {
#pragma unroll
        for (int bt = 0; bt < BT; ++bt) {
            if (b0 + bt >= B) break;
            float sm = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) sm += red[(w * BT + bt) * 256 + tid];
            part[((size_t)blockIdx.y * B + b0 + bt) * N + n] = sm;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// dec_attn: one CTA per (head, sequence), 128 threads.                     HF: modeling_gpt2.py:185-220 with layer_past
//   q,k,v[0..63] = bf16(bias + sum_s part[s][b][{0,d,2d} + 64 h + j]) ; k,v -> cache row `slot` ; scores over rows 0..slot (the causal
//   mask of a single query is "everything cached"), softmax in fp32, probabilities rounded to bf16 for P.V, fp32 normaliser.
// cache: [2 (k|v)][B][H][T_max][64] bf16 for this layer.
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dec_attn_kernel(const float* __restrict__ part, const float* __restrict__ bias, int S, int B, int d, int H,
                                                       bf16* __restrict__ kcache, bf16* __restrict__ vcache, int T_max,
                                                       const int32_t* __restrict__ slot_p, bf16* __restrict__ att16) {
    TTTS_DYN_SMEM(float, dec_smem);
    float* sc = dec_smem;                       // [T_max] scores -> probabilities
    __shared__ float qs[64];
    __shared__ float red[4];
    __shared__ float osum[4][64];
    pdl_launch_dependents();
    pdl_wait();
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = *slot_p;
    if (slot >= T_max) return;
    bf16* kc = kcache + ((size_t)b * H + h) * T_max * 64;
    bf16* vc = vcache + ((size_t)b * H + h) * T_max * 64;
    if (tid < 64) {
        float q = 0.f, k = 0.f, v = 0.f;
        const size_t col = (size_t)h * 64 + tid;
        float pq[DEC_SPLIT_MAX], pk[DEC_SPLIT_MAX], pv[DEC_SPLIT_MAX];
#pragma unroll
        for (int s = 0; s < DEC_SPLIT_MAX; ++s) {
            const float* p = part + ((size_t)(s < S ? s : 0) * B + b) * 3 * d;
            pq[s] = s < S ? p[col] : 0.f; pk[s] = s < S ? p[d + col] : 0.f; pv[s] = s < S ? p[2 * d + col] : 0.f;
        }
#pragma unroll
        for (int s = 0; s < DEC_SPLIT_MAX; ++s) { q += pq[s]; k += pk[s]; v += pv[s]; }
        q = bf16_round(q + bias[col]);
        kc[(size_t)slot * 64 + tid] = __float2bfloat16_rn(k + bias[d + col]);
        vc[(size_t)slot * 64 + tid] = __float2bfloat16_rn(v + bias[2 * d + col]);
        qs[tid] = q * 0.125f;                   // 64^-0.5, exact
    }
    __syncthreads();                            // this CTA's cache row and q are visible to all of its threads
    const int n_keys = slot + 1;
    // ---- scores: one key per thread ----
    float mx = -INFINITY;
    for (int t = tid; t < n_keys; t += 128) {
        const uint4* kr = reinterpret_cast<const uint4*>(kc + (size_t)t * 64);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 w = kr[c];
            s = fmaf(qs[8 * c + 0], bf16_lo(w.x), s); s = fmaf(qs[8 * c + 1], bf16_hi(w.x), s);
            s = fmaf(qs[8 * c + 2], bf16_lo(w.y), s); s = fmaf(qs[8 * c + 3], bf16_hi(w.y), s);
            s = fmaf(qs[8 * c + 4], bf16_lo(w.z), s); s = fmaf(qs[8 * c + 5], bf16_hi(w.z), s);
            s = fmaf(qs[8 * c + 6], bf16_lo(w.w), s); s = fmaf(qs[8 * c + 7], bf16_hi(w.w), s);
        }
        sc[t] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float l = 0.f;
    for (int t = tid; t < n_keys; t += 128) {
        const float p = __expf(sc[t] - mx);
        l += p;
        sc[t] = bf16_round(p);
    }
    l = warp_sum(l);
    if (lane == 0) red[warp] = l;
    __syncthreads();                            // also publishes sc[] for the second pass
    l = (red[0] + red[1]) + (red[2] + red[3]);
    // ---- P.V: lane owns dims 2 lane, 2 lane + 1 ; warp w owns keys w, w + 4, ... ----
    float o0 = 0.f, o1 = 0.f;
    for (int t0 = warp; t0 < n_keys; t0 += 4 * 8) {                  // 8 value rows in flight per warp (one row per trip was one L2 round trip each)
        uint32_t wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = t0 + 4 * u;
            wv[u] = t < n_keys ? *reinterpret_cast<const uint32_t*>(vc + (size_t)t * 64 + 2 * lane) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int t = t0 + 4 * u;
            const float p = t < n_keys ? sc[t] : 0.f;
            o0 = fmaf(p, bf16_lo(wv[u]), o0);
            o1 = fmaf(p, bf16_hi(wv[u]), o1);
        }
    }
    osum[warp][2 * lane] = o0;
    osum[warp][2 * lane + 1] = o1;
    __syncthreads();
    if (tid < 64) {
        const float o = ((osum[0][tid] + osum[1][tid]) + (osum[2][tid] + osum[3][tid])) / l;
        att16[(size_t)b * d + h * 64 + tid] = __float2bfloat16_rn(o);
    }
}

// dec_act: act16[b][n] = gelu_new(bf16(bias[n] + sum_s part[s][b][n]))  in packed bf16x2 like the c_fc GEMM epilogue.  2 columns per thread.
__global__ void __launch_bounds__(256) dec_act_kernel(const float* __restrict__ part, const float* __restrict__ bias, int S, int B, int N,
                                                      bf16* __restrict__ act16) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int n = (blockIdx.x * 256 + threadIdx.x) * 2;
    if (n >= N) return;
    float a0 = 0.f, a1 = 0.f;
    float2 ps[DEC_SPLIT_MAX];
#pragma unroll
    for (int s = 0; s < DEC_SPLIT_MAX; ++s)
        ps[s] = s < S ? *reinterpret_cast<const float2*>(part + ((size_t)s * B + b) * N + n) : make_float2(0.f, 0.f);
#pragma unroll
    for (int s = 0; s < DEC_SPLIT_MAX; ++s) { a0 += ps[s].x; a1 += ps[s].y; }
    const uint32_t pre = pack_bf16(a0 + bias[n], a1 + bias[n + 1]);
    *reinterpret_cast<uint32_t*>(act16 + (size_t)b * N + n) = gelu_new_bf2(pre);
}

// dec_head: logits[b][v] = bf16(bias[v] + sum_k enc16[b][k] * Wh[v][k])     nn.Linear layout [V][d]; ttts/gpt/model.py:432-437 (mel_head as lm_head, :366-371)
// one warp per vocabulary row (its d bf16 weights stay in registers for all sequences), 8 rows per CTA.
__global__ void __launch_bounds__(256) dec_head_kernel(const bf16* __restrict__ enc16, const bf16* __restrict__ Wh, const float* __restrict__ bias,
                                                       int B, int d, int V, float* __restrict__ logits) {
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    uint4 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                   // weights are requested ahead of the dependency wait (see dec_gemv)
        const int c = (i * 32 + lane) * 8;
        w[i] = (v < V && c < d) ? __ldg(reinterpret_cast<const uint4*>(Wh + (size_t)v * d + c)) : make_uint4(0, 0, 0, 0);
    }
    pdl_launch_dependents();
    pdl_wait();
    if (v >= V) return;
    const float bv = bias[v];
    for (int b = 0; b < B; ++b) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = (i * 32 + lane) * 8;
            if (c < d) {
                const uint4 x = *reinterpret_cast<const uint4*>(enc16 + (size_t)b * d + c);
                s = fmaf(bf16_lo(x.x), bf16_lo(w[i].x), s); s = fmaf(bf16_hi(x.x), bf16_hi(w[i].x), s);
                s = fmaf(bf16_lo(x.y), bf16_lo(w[i].y), s); s = fmaf(bf16_hi(x.y), bf16_hi(w[i].y), s);
                s = fmaf(bf16_lo(x.z), bf16_lo(w[i].z), s); s = fmaf(bf16_hi(x.z), bf16_hi(w[i].z), s);
                s = fmaf(bf16_lo(x.w), bf16_lo(w[i].w), s); s = fmaf(bf16_hi(x.w), bf16_hi(w[i].w), s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) logits[(size_t)b * V + v] = bf16_round(s + bv);
    }
}

__global__ void dec_advance_kernel(int32_t* slot) {
    pdl_wait();
    *slot += 1;
}

// cache fill from a ttts_gpt_forward(save_acts = 1) pass: rows t < n_pos of the packed c_attn output [B*T][3d] of one layer -> [2][B][H][T_max][64]
__global__ void __launch_bounds__(128) dec_kv_fill_kernel(const bf16* __restrict__ qkv, int T, int d, int H, int B, bf16* __restrict__ kcache,
                                                          bf16* __restrict__ vcache, int T_max) {
    const int t = blockIdx.x, b = blockIdx.y;
    const bf16* row = qkv + ((size_t)b * T + t) * 3 * d;
    for (int i = threadIdx.x; i < d / 8; i += 128) {                 // 8 elements (16 bytes) per access; a head is 8 accesses
        const int h = i >> 3, c = (i & 7) * 8;
        const size_t dst = (((size_t)b * H + h) * T_max + t) * 64 + c;
        *reinterpret_cast<uint4*>(kcache + dst) = *reinterpret_cast<const uint4*>(row + d + i * 8);
        *reinterpret_cast<uint4*>(vcache + dst) = *reinterpret_cast<const uint4*>(row + 2 * d + i * 8);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------------
static inline int64_t up256(int64_t n) { return (n + 255) / 256 * 256; }

struct DecWorkspace { int64_t resid, resid2, xn16, att16, act16, part, part2, total; };
static DecWorkspace dec_carve(int B, int d) {
    DecWorkspace w;
    int64_t o = 0;
    auto put = [&](int64_t bytes) { int64_t r = o; o += up256(bytes); return r; };
    w.resid = put((int64_t)B * d * 4);
    w.resid2 = put((int64_t)B * d * 4);
    w.xn16 = put((int64_t)B * d * 2);
    w.att16 = put((int64_t)B * d * 2);
    w.act16 = put((int64_t)B * 4 * d * 2);
    w.part = put((int64_t)DEC_SPLIT_MAX * B * 4 * d * 4);
    w.part2 = put((int64_t)DEC_SPLIT_MAX * B * 4 * d * 4);
    w.total = o;
    return w;
}
int64_t gpt_decode_workspace_bytes(int B, int d) { return dec_carve(B, d).total; }
int64_t gpt_kv_bytes(int layers, int B, int H, int T_max) { return (int64_t)layers * 2 * B * H * T_max * 64 * 2; }

// K slices: enough CTAs to cover the SMs, slices of >= 32 rows (4 per warp), at most DEC_SPLIT_MAX
static int dec_pick_split(int K, int N, int B) {
    const int cx = (N + 255) / 256, cz = (B + DEC_BT - 1) / DEC_BT;
    int S = 1;
    while (S < DEC_SPLIT_MAX && cx * S * cz < num_sms() && K / (S * 2) >= 32 && K % (S * 2) == 0) S *= 2;
    while (S < DEC_SPLIT_MAX && K / S > 1024 && K % (S * 2) == 0) S *= 2;           // the x slice [BT][K / S] must fit in shared memory
    return S;
}

static int dec_gemv(const bf16* x16, const bf16* W, float* part, int K, int N, int B, int* S_out, cudaStream_t st) {
    const int S = dec_pick_split(K, N, B);
    const int ks = K / S;
    TTTS_CHECK_ARG(K % S == 0 && N % 8 == 0, "decode: gemv shape K=%d N=%d", K, N);
    const dim3 grid((N + 255) / 256, S, (B + DEC_BT - 1) / DEC_BT);
    const size_t smem = ((size_t)DEC_BT * ks + 8 * DEC_BT * 256) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 48 * 1024, "decode: gemv slice of %d rows needs %zu bytes of shared memory", ks, smem);
    TTTS_CUDA(launch_pdl(dec_gemv_kernel<DEC_BT>, grid, dim3(256), smem, st, x16, W, part, K, N, B, ks));
    TTTS_LAUNCH_CHECK("dec_gemv");
    *S_out = S;
    return TTTS_OK;
}

template <int MODE>
static int dec_gemv_fused(const GemvPro& g, const bf16* W, float* part, int K, int N, int B, int* S_out, cudaStream_t st) {
    const int S = dec_pick_split(K, N, B);
    const int ks = K / S;
    TTTS_CHECK_ARG(K % S == 0 && N % 8 == 0 && ks % 2 == 0 && (MODE != 1 || K <= 1024), "decode: fused gemv shape K=%d N=%d", K, N);
    TTTS_CHECK_ARG(g.part_in != part, "decode: fused gemv reads and writes the same partial buffer");
    const dim3 grid((N + 255) / 256, S, (B + DEC_BT - 1) / DEC_BT);
    const size_t smem = ((size_t)DEC_BT * ks + 8 * DEC_BT * 256) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 48 * 1024, "decode: gemv slice of %d rows needs %zu bytes of shared memory", ks, smem);
    TTTS_CUDA(launch_pdl(dec_gemv_fused_kernel<DEC_BT, MODE>, grid, dim3(256), smem, st, g, W, part, K, N, B, ks));
    TTTS_LAUNCH_CHECK("dec_gemv_fused");
    *S_out = S;
    return TTTS_OK;
}

int gpt_kv_fill_layer(const bf16* qkv, int B, int T, int d, int H, int n_pos, bf16* kcache, bf16* vcache, int T_max, cudaStream_t st) {
    TTTS_CHECK_ARG(n_pos >= 1 && n_pos <= T && n_pos <= T_max, "decode: prefill of %d positions (sequence %d, cache %d)", n_pos, T, T_max);
    TTTS_CUDA(launch_plain(dec_kv_fill_kernel, dim3(n_pos, B), dim3(128), 0, st, qkv, T, d, H, B, kcache, vcache, T_max));
    TTTS_LAUNCH_CHECK("dec_kv_fill");
    return TTTS_OK;
}

int gpt_decode_step(const ttts_gpt_decode* a, cudaStream_t st) {
    TTTS_CHECK_ARG(a != nullptr, "decode: null args");
    const ttts_gpt_config& c = a->cfg;
    const int B = a->B, d = c.model_dim, H = c.heads, L = c.layers, T_max = a->T_max, Vm = c.n_mel_vocab;
    TTTS_CHECK_ARG(gpt_param_off(c, TTTS_P_TEXT_EMB, 0) == 0, "decode: bad GPT config");
    TTTS_CHECK_ARG(B >= 1 && B <= 1024, "decode: batch %d outside [1, 1024]", B);
    TTTS_CHECK_ARG(T_max >= 2 && T_max <= 8192, "decode: cache capacity %d outside [2, 8192]", T_max);
    TTTS_CHECK_ARG(a->text_positions >= 2 && a->text_positions < T_max, "decode: text segment of %d positions", a->text_positions);
    TTTS_CHECK_ARG(a->pos_shift == 0 || a->pos_shift == 1, "decode: pos_shift %d", a->pos_shift);
    TTTS_CHECK_ARG(a->codes && a->slot && a->params && a->params16 && a->kv && a->workspace && a->logits, "decode: null buffer");
    const DecWorkspace w = dec_carve(B, d);
    TTTS_CHECK_ARG(a->workspace_bytes >= w.total, "decode: workspace too small (%lld < %lld)", (long long)a->workspace_bytes, (long long)w.total);
    TTTS_CHECK_ARG(a->kv_bytes >= gpt_kv_bytes(L, B, H, T_max), "decode: cache too small (%lld < %lld)", (long long)a->kv_bytes,
                   (long long)gpt_kv_bytes(L, B, H, T_max));
    TTTS_CHECK_ARG(((uintptr_t)a->workspace & 255) == 0 && ((uintptr_t)a->kv & 15) == 0, "decode: workspace / cache alignment");
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    const float* p32 = a->params;
    const bf16* p16 = reinterpret_cast<const bf16*>(a->params16);
    float* resid = reinterpret_cast<float*>(ws + w.resid);
    bf16* xn16 = reinterpret_cast<bf16*>(ws + w.xn16);
    bf16* att16 = reinterpret_cast<bf16*>(ws + w.att16);
    bf16* act16 = reinterpret_cast<bf16*>(ws + w.act16);
    float* part = reinterpret_cast<float*>(ws + w.part);
    bf16* kv = reinterpret_cast<bf16*>(a->kv);
    const size_t kv_half = (size_t)B * H * T_max * 64;
    auto P = [&](int t, int l) { return gpt_param_off(c, t, l); };

    // fused schedule by default for B <= DEC_BT sequences (r2u: 1.08 -> 0.88 ms per code at B = 1; at B = 8 every CTA repeats the LayerNorm of four
    // sequences and the step is 3 % slower than with the separate row-op); TTTS_DECODE_FUSE=0 / 1 forces either
    static int fuse_env = -2;
    if (fuse_env == -2) { const char* e = getenv("TTTS_DECODE_FUSE"); fuse_env = !e ? -1 : (e[0] == '0' ? 0 : 1); }
    const int fuse = fuse_env >= 0 ? fuse_env : (B <= DEC_BT ? 1 : 0);
    float* R[2] = {resid, reinterpret_cast<float*>(ws + w.resid2)};
    float* PB[2] = {part, reinterpret_cast<float*>(ws + w.part2)};
    int cur = 0;                                                      // R[cur] holds the residual stream

    RowopArgs r;
    memset(&r, 0, sizeof(r));
    r.d = d; r.B = B; r.resid = resid; r.xn16 = xn16; r.part = part;
    r.codes = a->codes; r.ld_codes = a->ld_codes; r.slot = a->slot; r.text_positions = a->text_positions; r.pos_shift = a->pos_shift;
    r.T_max = T_max; r.Vm = Vm; r.n_pos_rows = c.max_mel_tokens + 2;
    r.Em = p32 + P(TTTS_P_MEL_EMB, 0); r.Pm = p32 + P(TTTS_P_MEL_POS, 0);
    int S = 1;
    float* last_part = part;
    for (int l = 0; l < L; ++l) {
        bf16* kc = kv + (size_t)l * 2 * kv_half;
        if (fuse) {
            // 5 launches per layer: [ln_1 +] c_attn | attention | attn c_proj | [residual + ln_2 +] c_fc | [gelu +] mlp c_proj
            if (l == 0) {
                r.mode = 0; r.S = 1; r.bias = nullptr; r.resid = R[cur];
                r.w1 = p32 + P(TTTS_P_LN1_W, l); r.b1 = p32 + P(TTTS_P_LN1_B, l); r.w2 = nullptr; r.b2 = nullptr;
                TTTS_CUDA(launch_pdl(dec_rowop_kernel, dim3(B), dim3(256), 0, st, r));
                TTTS_LAUNCH_CHECK("dec_rowop");
                TTTS_RUN(dec_gemv(xn16, p16 + P(TTTS_P_ATTN_W, l), PB[0], d, 3 * d, B, &S, st));
            } else {
                GemvPro g = {PB[1], S, p32 + P(TTTS_P_PR_B, l - 1), R[cur], R[cur ^ 1], p32 + P(TTTS_P_LN1_W, l), p32 + P(TTTS_P_LN1_B, l)};
                TTTS_RUN(dec_gemv_fused<1>(g, p16 + P(TTTS_P_ATTN_W, l), PB[0], d, 3 * d, B, &S, st));
                cur ^= 1;
            }
            TTTS_CUDA(launch_pdl(dec_attn_kernel, dim3(H, B), dim3(128), (size_t)T_max * sizeof(float), st, (const float*)PB[0],
                                 p32 + P(TTTS_P_ATTN_B, l), S, B, d, H, kc, kc + kv_half, T_max, (const int32_t*)a->slot, att16));
            TTTS_LAUNCH_CHECK("dec_attn");
            TTTS_RUN(dec_gemv(att16, p16 + P(TTTS_P_PROJ_W, l), PB[1], d, d, B, &S, st));
            {
                GemvPro g = {PB[1], S, p32 + P(TTTS_P_PROJ_B, l), R[cur], R[cur ^ 1], p32 + P(TTTS_P_LN2_W, l), p32 + P(TTTS_P_LN2_B, l)};
                TTTS_RUN(dec_gemv_fused<1>(g, p16 + P(TTTS_P_FC_W, l), PB[0], d, 4 * d, B, &S, st));
                cur ^= 1;
            }
            {
                GemvPro g = {PB[0], S, p32 + P(TTTS_P_FC_B, l), nullptr, nullptr, nullptr, nullptr};
                TTTS_RUN(dec_gemv_fused<2>(g, p16 + P(TTTS_P_PR_W, l), PB[1], 4 * d, d, B, &S, st));
            }
            last_part = PB[1];
            continue;
        }
        // embed (layer 0) or residual add of the previous layer's mlp c_proj ; ln_1
        r.mode = l == 0 ? 0 : 1; r.S = S; r.bias = l == 0 ? nullptr : p32 + P(TTTS_P_PR_B, l - 1);
        r.w1 = p32 + P(TTTS_P_LN1_W, l); r.b1 = p32 + P(TTTS_P_LN1_B, l); r.w2 = nullptr; r.b2 = nullptr;
        TTTS_CUDA(launch_pdl(dec_rowop_kernel, dim3(B), dim3(256), 0, st, r));
        TTTS_LAUNCH_CHECK("dec_rowop");
        // c_attn
        TTTS_RUN(dec_gemv(xn16, p16 + P(TTTS_P_ATTN_W, l), part, d, 3 * d, B, &S, st));
        TTTS_CUDA(launch_pdl(dec_attn_kernel, dim3(H, B), dim3(128), (size_t)T_max * sizeof(float), st, (const float*)part,
                             p32 + P(TTTS_P_ATTN_B, l), S, B, d, H, kc, kc + kv_half, T_max, (const int32_t*)a->slot, att16));
        TTTS_LAUNCH_CHECK("dec_attn");
        // attn c_proj ; residual add ; ln_2
        TTTS_RUN(dec_gemv(att16, p16 + P(TTTS_P_PROJ_W, l), part, d, d, B, &S, st));
        r.mode = 1; r.S = S; r.bias = p32 + P(TTTS_P_PROJ_B, l);
        r.w1 = p32 + P(TTTS_P_LN2_W, l); r.b1 = p32 + P(TTTS_P_LN2_B, l);
        TTTS_CUDA(launch_pdl(dec_rowop_kernel, dim3(B), dim3(256), 0, st, r));
        TTTS_LAUNCH_CHECK("dec_rowop");
        // c_fc ; gelu_new ; mlp c_proj
        TTTS_RUN(dec_gemv(xn16, p16 + P(TTTS_P_FC_W, l), part, d, 4 * d, B, &S, st));
        TTTS_CUDA(launch_pdl(dec_act_kernel, dim3((4 * d / 2 + 255) / 256, B), dim3(256), 0, st, (const float*)part, p32 + P(TTTS_P_FC_B, l), S, B,
                             4 * d, act16));
        TTTS_LAUNCH_CHECK("dec_act");
        TTTS_RUN(dec_gemv(act16, p16 + P(TTTS_P_PR_W, l), part, 4 * d, d, B, &S, st));
    }
    // last residual add ; ln_f ; final_norm ; mel head
    r.mode = 1; r.S = S; r.bias = p32 + P(TTTS_P_PR_B, L - 1); r.part = last_part; r.resid = R[cur];
    r.w1 = p32 + P(TTTS_P_LNF_W, 0); r.b1 = p32 + P(TTTS_P_LNF_B, 0); r.w2 = p32 + P(TTTS_P_FN_W, 0); r.b2 = p32 + P(TTTS_P_FN_B, 0);
    TTTS_CUDA(launch_pdl(dec_rowop_kernel, dim3(B), dim3(256), 0, st, r));
    TTTS_LAUNCH_CHECK("dec_rowop");
    TTTS_CUDA(launch_pdl(dec_head_kernel, dim3((Vm + 7) / 8), dim3(256), 0, st, (const bf16*)xn16, p16 + P(TTTS_P_MEL_HEAD_W, 0),
                         p32 + P(TTTS_P_MEL_HEAD_B, 0), B, d, Vm, a->logits));
    TTTS_LAUNCH_CHECK("dec_head");
    TTTS_CUDA(launch_pdl(dec_advance_kernel, dim3(1), dim3(1), 0, st, a->slot));
    TTTS_LAUNCH_CHECK("dec_advance");
    return TTTS_OK;
}

}  // namespace ttts
