/*
 * ttts_b200 -- C ABI of the Blackwell (sm_100a) hot-path library `libttts_b200.so`.
 *
 * The reference (adelacvg/ttts) is pure Python/PyTorch and has NO FFI of its own; the boundary a
 * maintainer binds is therefore the set of torch ops its nn.Modules dispatch on the hot path.
 * Each entry point below names the reference op site (file:line under /root/reference, or HF:
 * transformers 5.5 modeling_gpt2.py) whose arithmetic it replaces.  See INTEGRATION.md for the
 * ctypes stubs the reference side would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ttts_status otherwise; the text of the last
 *     error on the calling thread is available from ttts_last_error();
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch tensors); the library
 *     allocates no persistent device memory;
 *   - every call takes the cudaStream_t to launch on (as void*), performs no host sync;
 *   - bf16 = raw 16-bit brain float, row-major unless stated.
 */
#ifndef TTTS_B200_H
#define TTTS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TTTS_OK = 0,
    TTTS_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    TTTS_ERR_CUDA = -2,      /* CUDA runtime / driver error      */
    TTTS_ERR_ARCH = -3       /* device is not sm_100             */
} ttts_status;

int ttts_version(void);
const char* ttts_last_error(void);
/* 1 if the current device can run the sm_100a kernels */
int ttts_device_ok(void);
/* number of kernels this library has launched since it was loaded (bench.py's gpu_launches) */
unsigned long long ttts_launch_count(void);
/* bracket every launch of ONE kernel family with CUDA events on its stream (bench.py's live roofline measurement):
 * on = 1 tcgen05 GEMMs | 2 conv1d input-gradient | 3 conv1d weight-gradient | 5 conv1d_tcs | 0 off.  The read-out sums ms and
 * algorithmic FLOPs (2 M N K; convolutions: 2 B Tout Cin Cout K) of the bracketed launches since the enable. */
void ttts_prof_gemm_enable(int on);
int ttts_prof_gemm_read(double* ms_total, double* flops_total, long long* launches);

/* --------------------------------------------------------------------------------------------
 * GEMM on the 5th-gen tensor cores (tcgen05.mma, TMA-fed, TMEM accumulators).
 *   D[M,N] = A[M,K] * B[K,N]   (bf16 x bf16 -> fp32 accumulate) followed by a fused epilogue.
 * Replaces the cuBLAS addmm / linear calls of HF Conv1D (HF: pytorch_utils.py:119-123) and
 * nn.Linear heads (ttts/gpt/model.py:348-349,432-437) and their autograd backward.
 *   a_mn = 0: A stored [M][K] (K contiguous, row stride lda);  a_mn = 1: A stored [K][M] (M contiguous)
 *   b_mn = 0: B stored [N][K] (K contiguous, row stride ldb);  b_mn = 1: B stored [K][N] (N contiguous)
 * ------------------------------------------------------------------------------------------ */
typedef enum {
    TTTS_EPI_BF16 = 0,       /* out_bf16 = acc (+bias[n])                                              */
    TTTS_EPI_GELU = 1,       /* aux_out_bf16 = pre = bf16(acc+bias) ; out_bf16 = gelu_new(pre)         */
    TTTS_EPI_RESID = 2,      /* out_f32 = resid_f32 + dropout(bf16(acc+bias))   (out may alias resid)  */
    TTTS_EPI_DGELU = 3,      /* out_bf16 = acc * gelu_new'(aux_bf16[m,n])                              */
    TTTS_EPI_F32_ADD = 4,    /* out_f32 += acc   (red.global.add; used by split-K weight gradients)    */
    TTTS_EPI_F32 = 5         /* out_f32 = acc (+bias)                                                  */
} ttts_epilogue;

typedef struct {
    int32_t M, N, K;
    const void* A; int32_t lda; int32_t a_mn;
    const void* B; int32_t ldb; int32_t b_mn;
    int32_t epi;
    void* out; int32_t ldo;
    const float* bias;          /* [N] or NULL */
    const void* aux; int32_t ldaux;   /* RESID: fp32 residual in ; DGELU: bf16 pre-activation */
    void* aux_out; int32_t ldaux_out; /* GELU: bf16 pre-activation out (may be NULL) */
    int32_t split_k;            /* >=1; only with TTTS_EPI_F32_ADD */
    /* dropout for TTTS_EPI_RESID (p = drop_thresh16 / 65536; 0 disables) */
    uint32_t drop_thresh16; float drop_scale; uint64_t drop_seed;
} ttts_gemm_args;

int ttts_gemm_bf16(const ttts_gemm_args* args, void* stream);


/* --------------------------------------------------------------------------------------------
 * UnifiedVoice GPT train step (ttts/gpt/model.py:453-510 driven by ttts/gpt/train.py:99-121).
 *
 * Parameters live in ONE flat fp32 buffer (plus a bf16 shadow at the same element offsets that feeds the
 * tensor cores, and a flat fp32 gradient buffer laid out identically so that data-parallel training is a
 * single NCCL all-reduce).  ttts_gpt_param_offset() gives the element offset of every reference
 * state_dict tensor inside those buffers; the Python module exposes them as nn.Parameter
 * views with the reference's names and shapes.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t layers, model_dim, heads;
    int32_t max_text_tokens, max_mel_tokens;     /* position tables hold max+2 rows (gpt/model.py:339)   */
    int32_t n_text_vocab, n_mel_vocab;           /* 257 = number_text_tokens*types+1 ; 1026               */
    int32_t start_text_token, stop_text_token;   /* 255, 0                                                */
    int32_t start_mel_token, stop_mel_token;     /* 1024, 1025                                            */
    int32_t mel_length_compression;              /* 1024                                                  */
} ttts_gpt_config;

typedef enum {
    TTTS_P_TEXT_EMB = 0, TTTS_P_MEL_EMB, TTTS_P_TEXT_POS, TTTS_P_MEL_POS,
    /* per layer */
    TTTS_P_LN1_W, TTTS_P_LN1_B, TTTS_P_ATTN_W, TTTS_P_ATTN_B, TTTS_P_PROJ_W, TTTS_P_PROJ_B,
    TTTS_P_LN2_W, TTTS_P_LN2_B, TTTS_P_FC_W, TTTS_P_FC_B, TTTS_P_PR_W, TTTS_P_PR_B,
    /* top */
    TTTS_P_LNF_W, TTTS_P_LNF_B, TTTS_P_FN_W, TTTS_P_FN_B,
    TTTS_P_TEXT_HEAD_W, TTTS_P_TEXT_HEAD_B, TTTS_P_MEL_HEAD_W, TTTS_P_MEL_HEAD_B,
    TTTS_P_COUNT
} ttts_gpt_tensor;

/* element offset of a tensor in the flat buffers (layer ignored for non-layer tensors); <0 on error */
int64_t ttts_gpt_param_offset(const ttts_gpt_config* cfg, int32_t tensor, int32_t layer);
/* number of elements of that tensor */
int64_t ttts_gpt_param_numel(const ttts_gpt_config* cfg, int32_t tensor);
/* total elements of the flat buffers (every tensor padded to a multiple of 64 elements) */
int64_t ttts_gpt_param_count(const ttts_gpt_config* cfg);
/* [begin,end) element range of backward stage s (0 = heads+final norms, 1..L = layer L-s, L+1 = embeddings):
 * the gradients of that range are final once ttts_gpt_backward has run stages 0..s. */
int32_t ttts_gpt_stage_range(const ttts_gpt_config* cfg, int32_t stage, int64_t* begin, int64_t* end);

/* activations / scratch; save_acts=1 keeps what backward needs (training), 0 reuses one layer's buffers */
int64_t ttts_gpt_workspace_bytes(const ttts_gpt_config* cfg, int32_t B, int32_t TL, int32_t CL, int32_t save_acts);

typedef enum {
    TTTS_WS_MEL_LOGITS = 0,   /* bf16 [B*(CL+2), ld_mel]   (ld via ttts_gpt_logits_ld)   */
    TTTS_WS_TEXT_LOGITS = 1,  /* bf16 [B*(TL+2), ld_text]                                 */
    TTTS_WS_LATENT = 2,       /* fp32 [B*T, d] final_norm output (return_latent)          */
    TTTS_WS_RESID = 3,        /* fp32 [B*T, d] residual stream entering layer `layer` (layer==L: input of ln_f) */
    TTTS_WS_TOKENS = 4        /* int32 text_in | text_tgt | mel_in | mel_tgt              */
} ttts_gpt_ws_item;
int64_t ttts_gpt_workspace_offset(const ttts_gpt_config* cfg, int32_t B, int32_t TL, int32_t CL, int32_t save_acts,
                                  int32_t item, int32_t layer);   /* byte offset, <0 on error */
int32_t ttts_gpt_logits_ld(int32_t vocab);

typedef struct {
    ttts_gpt_config cfg;
    int32_t B, TL, CL;                 /* (clipped) text / code lengths; T = TL + CL + 4                        */
    const int64_t* text; int32_t ld_text;        /* [B, >=TL] raw text tokens                                     */
    int64_t* codes; int32_t ld_codes;            /* [B, >=CL] raw mel codes -- mutated in place (set_mel_padding) */
    const int64_t* wav_lengths;                  /* [B]                                                            */
    const float* params;               /* flat fp32 parameters                                                   */
    const void* params16;              /* flat bf16 shadow (ttts_cast_bf16 or ttts_adamw_step keep it fresh)     */
    float* grads;                      /* flat fp32 gradients (backward ACCUMULATES)                             */
    void* workspace; int64_t workspace_bytes;
    float* losses;                     /* device [2]: loss_text, loss_mel                                        */
    int32_t save_acts;                 /* 1: keep activations for backward                                       */
    int32_t want_latent;               /* 1: also write TTTS_WS_LATENT and skip heads/losses                     */
    float drop_p; uint64_t seed;       /* dropout (embd, attn-prob, attn-out, mlp-out); 0 = eval                 */
    /* backward only */
    const float* gscale_text; const float* gscale_mel;  /* device scalars dL/dloss_* (NULL = 1)                   */
    float weight_text, weight_mel;                       /* host-side multipliers of the two losses               */
} ttts_gpt_io;

int ttts_gpt_forward(const ttts_gpt_io* io, void* stream);
/* runs backward stages [stage_begin, stage_end) ; see ttts_gpt_stage_range */
int ttts_gpt_backward(const ttts_gpt_io* io, int32_t stage_begin, int32_t stage_end, void* stream);

/* --------------------------------------------------------------------------------------------
 * KV-cache decode: autoregressive code generation, one new code per sequence per call -- the cached branch of
 * GPT2InferenceModel.forward (ttts/gpt/model.py:106-171 with past_key_values, :144-147 single-token embedding) as
 * UnifiedVoice.inference_speech drives it (:533-562).  The cache holds, per layer, K and V of every position so far:
 * bf16 [layers][2 (k|v)][B][heads][T_max][64].  Fill it from a prompt with ttts_gpt_forward(save_acts = 1) followed by
 * ttts_gpt_kv_prefill, then call ttts_gpt_decode_step once per generated code.
 * ------------------------------------------------------------------------------------------ */
int64_t ttts_gpt_kv_bytes(const ttts_gpt_config* cfg, int32_t B, int32_t T_max);
int64_t ttts_gpt_decode_workspace_bytes(const ttts_gpt_config* cfg, int32_t B);
/* copy K / V of sequence positions [0, n_pos) of every layer out of the workspace of a ttts_gpt_forward(save_acts = 1) pass */
int ttts_gpt_kv_prefill(const ttts_gpt_io* io, void* kv, int64_t kv_bytes, int32_t T_max, int32_t n_pos, void* stream);

typedef struct {
    ttts_gpt_config cfg;
    int32_t B, T_max;                  /* sequences ; cache capacity in positions                                          */
    int32_t text_positions;            /* TL + 2: cache slots taken by [start, text..., stop]                              */
    int32_t pos_shift;                 /* 0: mel position index = the uncached path's (model.py:134-142) ;
                                          1: the index the reference's cached branch computes (model.py:144-147), one later */
    const int64_t* codes; int32_t ld_codes;   /* [B, >= n] codes so far; the token fed is codes[b, slot - text_positions - 1]   */
    int32_t* slot;                     /* device scalar: cache slot of the token being fed (= positions cached so far);
                                          read by every kernel of the step and advanced by one at its end, so a captured
                                          step (CUDA graph) replays for the next position unchanged                        */
    const float* params; const void* params16;
    void* kv; int64_t kv_bytes;
    void* workspace; int64_t workspace_bytes;
    float* logits;                     /* out: fp32 [B, n_mel_vocab] mel-head logits of the fed token (bf16-rounded values) */
} ttts_gpt_decode;
int ttts_gpt_decode_step(const ttts_gpt_decode* args, void* stream);

/* --------------------------------------------------------------------------------------------
 * step tail: get_grad_norm + clip_grad_norm_(1.0) + AdamW (ttts/gpt/train.py:22-31,114-118)
 * ------------------------------------------------------------------------------------------ */
int ttts_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* norm_out[0] = ||g||_2 ; scratch: >= 1024 floats */
int ttts_grad_norm(const float* grads, int64_t n, float* scratch, float* norm_out, void* stream);
/* p, m, v updated in place; p16 (optional) receives the bf16 shadow.  grad is scaled by grad_scale (1/world for a
 * summed all-reduce) and clipped to max_norm using *norm (the norm of the UNSCALED buffer); max_norm<=0 disables. */
int ttts_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params16, int64_t n,
                    const float* norm, float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, void* stream);

/* --------------------------------------------------------------------------------------------
 * individual kernels (exported for the per-kernel parity tests)
 * ------------------------------------------------------------------------------------------ */
/* y = LN(x) (dbl=0) or LN2(LN1(x)) (dbl=1); stats: [M,2] or [M,4] (mean,rstd)            (HF:modeling_gpt2.py:273,304,628) */
int ttts_layernorm_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, void* y,
                       float* stats, int32_t M, int32_t d, int32_t dbl, int32_t out_bf16, void* stream);
int ttts_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, const float* stats, const float* w1, const float* b1,
                       const float* w2, const float* g_in, float* g_out, void* g16_out, float* dw1, float* db1, float* dw2,
                       float* db2, float* dbias_next, int32_t M, int32_t d, int32_t dbl, void* stream);
/* causal flash attention on the packed c_attn output [B*T, 3d] (HF:modeling_gpt2.py:144-226).
 * ttts_attn_bwd scratch: ttts_attn_bwd_scratch_floats(B,T,H) floats (row dots + the fp32 dQ accumulator). */
int64_t ttts_attn_bwd_scratch_floats(int32_t B, int32_t T, int32_t H);
int ttts_attn_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t T, int32_t H, float drop_p, uint64_t seed, void* stream);
int ttts_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta_scratch, void* dqkv,
                  int32_t B, int32_t T, int32_t H, float drop_p, uint64_t seed, void* stream);
/* keep mask [BH, T, T] (1 = kept) of the attention-probability dropout for (drop_p, seed): with it torch can reproduce
 * ttts_attn_fwd/bwd under dropout exactly (HF:modeling_gpt2.py:216 attn_dropout; the hash is ours) */
int ttts_attn_dropout_mask(uint8_t* mask, int32_t BH, int32_t T, float drop_p, uint64_t seed, void* stream);
/* keep mask one dropout site of ttts_gpt_forward(io) draws for (io->drop_p, io->seed): site 0 = embedding (HF drop, modeling_gpt2.py
 * GPT2Model.forward), 1 = attention probabilities (attn_dropout; mask [rows = B*H, cols = T, T]), 2 = attention output / 3 = MLP output
 * (resid_dropout; mask [rows = B*T, cols = d]); `layer` is ignored for site 0.  Kept elements are scaled by 32768 / (32768 - round(p * 32768)).
 * With the four masks a plain-torch restatement reproduces the TRAINING-mode step exactly (tests/test_gpu_gpt.py). */
int ttts_gpt_dropout_mask(uint8_t* mask, int32_t site, int32_t layer, int32_t rows, int32_t cols, float drop_p, uint64_t seed, void* stream);
/* mean cross-entropy over rows of bf16 logits [rows, ld] (ttts/gpt/model.py:508-509) */
int ttts_ce_fwd(const void* logits, int32_t ld, int32_t V, const int32_t* targets, int32_t rows, float* row_loss, float* row_lse,
                float* loss_out, void* stream);
int ttts_ce_bwd(const void* logits, int32_t ld, int32_t V, const int32_t* targets, int32_t rows, const float* row_lse,
                const float* gscale, float weight, void* dlogits, void* stream);

/* --------------------------------------------------------------------------------------------
 * VQ-VAE encode front end
 * ------------------------------------------------------------------------------------------ */
/* RVQ (n_q=1) lookup: EuclideanCodebook.quantize/forward, ttts/vqvae/core_vq.py:174-230, VectorQuantization.forward 303-322.
 *   x: fp32, layout_bdn=1 -> [B, D, Nn] (the module's "b d n"), vector v = b*Nn+n ; layout_bdn=0 -> [B, D] rows (Nn ignored)
 *   codes: int64 [N] ; quantized (optional): same layout as x, = E[code] or the straight-through value x + (q - x)
 *   commit_out (optional): mean((q - x)^2) ; hist [K], embed_sum [K,D] (optional, must be zeroed): EMA statistics
 *   workspace: ttts_vq_workspace_floats(N, K) floats */
int64_t ttts_vq_workspace_floats(int32_t N, int32_t K);
int ttts_vq_forward(const float* x, int32_t B, int32_t D, int32_t Nn, int32_t layout_bdn, const float* embed, int32_t K, int64_t* codes,
                    float* quantized, int32_t straight_through, float* commit_out, float* hist, float* embed_sum, float* workspace,
                    void* stream);
/* cluster_size/embed_avg EMA + Laplace-smoothed renormalisation of embed (core_vq.py:217-228); scratch1: 1 float */
int ttts_vq_ema_update(float* embed, float* embed_avg, float* cluster_size, const float* hist, const float* embed_sum, int32_t K,
                       int32_t D, float decay, float eps, float* scratch1, void* stream);
/* dx = dquantized + dcommit * 2 (x - q) / (N*D)  (dquantized / dcommit may be NULL) */
int ttts_vq_backward(const float* x, int32_t B, int32_t D, int32_t Nn, int32_t layout_bdn, const float* embed, const int64_t* codes,
                     const float* dquantized, const float* dcommit, float* dx, void* stream);

/* Framed rFFT + magnitude (+ sparse mel + log) : spectrogram_torch / spec_to_mel_torch / mel_spectrogram_torch
 * (ttts/utils/data_utils.py:52-156) and MelSpectrogramFeatures (ttts/vocoder/feature_extractors.py:28-49).
 *   wav [B, L] fp32 ; frames are taken from the reflect-padded signal (pad samples each side), hop apart, n_fft long
 *   window [n_fft] ; twiddle [n_fft/2+1] complex (re,im) = exp(-2 pi i k / n_fft)
 *   spec_out (optional) [B, n_fft/2+1, F] = sqrt(re^2 + im^2 + eps_inside)
 *   mel_out (optional) [B, n_mels, F] = log(max(basis @ spec, log_floor)); the basis is sparse: band m covers bins
 *   band_lo[m] .. band_lo[m] + (band_off[m+1]-band_off[m]) with weights band_w[band_off[m] ..] */
int ttts_stft_mel(const float* wav, int32_t B, int32_t L, int32_t n_fft, int32_t hop, int32_t pad, const float* window,
                  const float* twiddle, float eps_inside, float* spec_out, int32_t n_mels, const int32_t* band_lo,
                  const int32_t* band_off, const float* band_w, float log_floor, float* mel_out, int32_t n_frames, void* stream);
/* backward of ttts_stft_mel's mel_out with respect to the waveform (the mel-reconstruction loss of the VQ-VAE-GAN step differentiates through
 * mel_spectrogram_torch(y_hat), ttts/vqvae/train.py:357-366,389; csrc/stft_bwd.cu; GPU-tested through the generator step since r2a).
 * dwav [B, L] ACCUMULATES (zero it first); dlogmel [B, n_mels, n_frames]. */
int ttts_stft_mel_bwd(const float* wav, int32_t B, int32_t L, int32_t n_fft, int32_t hop, int32_t pad, const float* window, float eps_inside,
                      int32_t n_mels, const int32_t* band_lo, const int32_t* band_off, const float* band_w, float log_floor,
                      const float* dlogmel, int32_t n_frames, float* dwav, void* stream);
int ttts_logmel(const float* spec, int32_t B, int32_t bins, int32_t F, int32_t n_mels, const int32_t* band_lo, const int32_t* band_off,
                const float* band_w, float log_floor, float* mel_out, void* stream);

/* fp32 Conv1d with the reference blocks' elementwise work fused (PosteriorAudioEncoder / WN / MelStyleEncoder,
 * ttts/vqvae/vq2.py:667-745, modules.py:136-318,560-566,686-764).  x [B,Cin,Tin], w [Cout,Cin,K] (torch layout), y [B,Cout',Tout]
 *   pre_lrelu: leaky_relu(0.1) on the input ; y = (post(conv + bias) + resid) * out_scale * mask ; accumulate: y += instead of =
 *   post: 0 none | 1 GLU (Cout = 2C -> C channels: a * sigmoid(b)) | 2 Mish | 3 WN gate tanh(a+cond_a) * sigmoid(b+cond_b) */
int ttts_conv1d_f32(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                    int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate,
                    const float* mask, int32_t post, const float* cond, int32_t cond_ld, void* stream);
/* the same convolution on the split-reduction kernel (groups = 2 or 4 warp groups share a tile's reduction; conv1d_split.cu), whatever
 * the layer: per-kernel parity tests.  ttts_conv1d_f32 itself picks it for latency-bound layers when TTTS_CONV_SPLIT=1. */
int ttts_conv1d_f32_split(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                          int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate,
                          const float* mask, int32_t post, const float* cond, int32_t cond_ld, int32_t groups, void* stream);
/* A stride-1 "same" convolution (pad = dil (K - 1) / 2; Cin % 8 == 0, 16 <= Cin <= 192; Cout in {32, 64, 96, 128, 192, 384}) of the encoder's
 * ResBlock1 / WN stacks (ttts/vqvae/modules.py:136-318) on the tcgen05 tensor cores with split-bf16 operands -- hi*hi + hi*lo + lo*hi,
 * fp32 accumulation in TMEM, ~1.5e-5 relative to fp32 -- and no im2col: a tap is a row shift of ONE channel-last window in shared memory
 * (csrc/conv1d_tcs.cu).  The weights are split once per layer into a caller-owned bf16 buffer of ttts_conv1d_tcs_weight_elems() elements:
 *   y = (post(conv(lrelu?(x), w) + bias) + resid) * out_scale * mask ; accumulate: y += ; post: 0 none | 3 WN gate (Cout = 384 -> 192 channels,
 *   tanh(a + cond_a) * sigmoid(g + cond_g), cond [B, 384] with row stride cond_ld, may be null) ;  flags bit 0: set the descriptor base offset (debug: measured WRONG on B200, the swizzle follows absolute address bits) */
int64_t ttts_conv1d_tcs_weight_elems(int32_t Cout, int32_t Cin, int32_t K);
int ttts_conv1d_tcs_prep_weights(const float* w, void* ws_bf16, int32_t Cout, int32_t Cin, int32_t K, void* stream);
int ttts_conv1d_tcs(const float* x, const void* ws_bf16, const float* bias, float* y, int32_t B, int32_t Cin, int32_t T, int32_t Cout, int32_t K,
                    int32_t dil, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate, const float* mask, int32_t post,
                    const float* cond, int32_t cond_ld, int32_t flags, void* stream);
/* Backward of that convolution (autograd of nn.Conv1d; next scope row, SURVEY.md 8f-1 -- first developed on the CPU
 * emulation of the source, GPU-tested against torch.autograd since r2a).  dy [B,Cout,Tout] is the gradient of the raw convolution output (before any fused post / residual).
 *   bwd_input : dx[B,Cin,Tin] (+)= lrelu'(x) * conv_transpose(dy, w)      x only read when pre_lrelu (the forward's input)
 *   bwd_weight: dw[Cout,Cin,K] += dy (*) lrelu(x) ; db[Cout] += sum dy (db may be NULL).  Gradients ACCUMULATE: zero them first. */
int ttts_conv1d_bwd_input(const float* dy, const float* w, const float* x, float* dx, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                          int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, int32_t accumulate, void* stream);
int ttts_conv1d_bwd_weight(const float* dy, const float* x, float* dw, float* db, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                           int32_t stride, int32_t dil, int32_t pad, int32_t pre_lrelu, void* stream);
/* db[C] += sum over (b, t) of dy[B,C,T].  ConvTranspose1d (the Generator's up-sampling, vq2.py:369-378) is served by these three entry points
 * with the roles swapped: forward = ttts_conv1d_bwd_input(x, w), dx = ttts_conv1d_f32(dy, w), dw = ttts_conv1d_bwd_weight(dy := x, x := dy). */
int ttts_bias_grad(const float* dy, float* db, int32_t B, int32_t C, int32_t T, void* stream);
/* torch weight_norm (dim 0): w[co,:] = g[co] * v[co,:] / ||v[co,:]|| */
int ttts_weight_norm(const float* v, const float* g, float* w, int32_t Cout, int32_t n_per_out, void* stream);
/* Activation1d(SnakeBeta(alpha_logscale)) : 2x kaiser-sinc upsample, x + sin^2(e^a x)/e^b, 2x low-pass downsample
 * (alias_free_torch/act.py:8-28, activations.py:62-119); filt12 = kaiser_sinc_filter1d(0.25, 0.3, 12) */
int ttts_snake_aa(const float* x, const float* log_alpha, const float* log_beta, const float* filt12, float* y, int32_t B, int32_t C, int32_t T,
                  void* stream);
/* MelStyleEncoder pieces: small masked multi-head attention over T <= 64 frames ([B,C,T] channel-major), masked temporal mean */
int ttts_mha_small(const float* q, const float* k, const float* v, const int64_t* lens, float* out, int32_t B, int32_t C, int32_t T, int32_t heads,
                   float temperature, void* stream);
int ttts_masked_mean(const float* x, const int64_t* lens, float* y, int32_t B, int32_t C, int32_t T, void* stream);
/* z = (m + eps * exp(logs)) * mask with stats = [B, 2C, T] (m | logs)  (vq2.py:742-744; eps NULL = 0) */
int ttts_posterior_sample(const float* stats, const float* eps, const float* mask, float* z, int32_t B, int32_t C, int32_t T, void* stream);

/* --------------------------------------------------------------------------------------------
 * Training kernels of the VQ-VAE encode half (next scope row; csrc/encoder_bwd.cu -- developed on the CPU emulation, GPU-tested since r2a,
 * validated on the CPU emulation of the source against tests/ref_kernels.py).  fp32, [B, C, T] channel-major.
 * ------------------------------------------------------------------------------------------ */
int ttts_ew_add(const float* a, const float* b, float* o, int64_t n, void* stream);
int ttts_ew_scale(const float* a, float s, float* o, int64_t n, void* stream);
int ttts_ew_mul_mask(const float* a, const float* mask, float* o, int32_t B, int32_t C, int32_t T, void* stream);            /* mask [B, T] */
/* Conv1dGLU gate (modules.py:560-566): raw [B,2C,T] = [a | g] ; backward = 0: out [B,C,T] = a sigmoid(g) ; 1: out = d raw from dy [B,C,T] */
int ttts_glu(const float* raw, const float* dy, float* out, int32_t B, int32_t C, int32_t T, int32_t backward, void* stream);
/* Mish: backward = 0: out = x tanh(softplus(x)) ; 1: out = dy * d/dx */
int ttts_mish(const float* x, const float* dy, float* out, int64_t n, int32_t backward, void* stream);
/* WN gate (modules.py:195-201): raw [B,2H,T], cond [B,2H] or NULL ; backward = 0: out [B,H,T] ; 1: out = d raw, dcond [B,2H] (or NULL) */
int ttts_wn_gate(const float* raw, const float* cond, const float* dy, float* out, float* dcond, int32_t B, int32_t H, int32_t T,
                 int32_t backward, void* stream);
/* leaky ReLU with an explicit slope / tanh: backward = 0: out = f(x) ; 1: out = dy f'(x) */
int ttts_lrelu(const float* x, const float* dy, float* out, int64_t n, float slope, int32_t backward, void* stream);
int ttts_tanh(const float* x, const float* dy, float* out, int64_t n, int32_t backward, void* stream);
/* out[b,c,t] = a[b,c,t] + v[c] (per_batch = 0) or v[b,c] (per_batch = 1) ; ttts_sum_t: out[row] = sum_t a[row, t] */
int ttts_add_bcast(const float* a, const float* v, float* out, int32_t B, int32_t C, int32_t T, int32_t per_batch, void* stream);
int ttts_sum_t(const float* a, float* out, int32_t rows, int32_t T, void* stream);
/* backward of ttts_weight_norm: dv [Cout, n], dg [Cout] from dw */
int ttts_weight_norm_bwd(const float* dw, const float* v, const float* g, float* dv, float* dg, int32_t Cout, int32_t n_per_out, void* stream);
/* backward of ttts_snake_aa: dx [B,C,T] ; dla / dlb [C] ACCUMULATE (zero them first) */
int ttts_snake_aa_bwd(const float* dy, const float* x, const float* log_alpha, const float* log_beta, const float* filt12, float* dx,
                      float* dla, float* dlb, int32_t B, int32_t C, int32_t T, void* stream);
/* backward of ttts_mha_small (T <= 64) */
int ttts_mha_small_bwd(const float* dout, const float* q, const float* k, const float* v, const int64_t* lens, float* dq, float* dk, float* dv,
                       int32_t B, int32_t C, int32_t T, int32_t heads, float temperature, void* stream);
/* backward of ttts_masked_mean: dy [B, C] -> dx [B, C, T] */
int ttts_masked_mean_bwd(const float* dy, const int64_t* lens, float* dx, int32_t B, int32_t C, int32_t T, void* stream);
/* backward of ttts_posterior_sample: dz [B,C,T] -> dstats [B,2C,T] */
int ttts_posterior_sample_bwd(const float* dz, const float* stats, const float* eps, const float* mask, float* dstats, int32_t B, int32_t C,
                              int32_t T, void* stream);

/* adversarial losses of the VQ-VAE-GAN step (ttts/vqvae/losses.py:7-44; csrc/gan_losses.cu, same validation status).  Deterministic
 * two-stage reductions; scratch = 256 floats; out / dL are device scalars. */
int ttts_lsgan_loss(const float* x, float c, int64_t n, float* scratch, float* out, void* stream);                 /* mean((c - x)^2)            */
int ttts_lsgan_loss_bwd(const float* x, float c, const float* dL, int64_t n, float* dx, void* stream);
int ttts_l1_mean(const float* a, const float* b, int64_t n, float* scratch, float* out, void* stream);               /* mean(|a - b|), a detached  */
int ttts_l1_mean_bwd(const float* a, const float* b, const float* dL, int64_t n, float* db, void* stream);
/* kl_loss (losses.py:47-61): out2[0] = sum((logs_p - logs_q - 0.5 + 0.5 (z_p - m_p)^2 exp(-2 logs_p)) mask) / sum(mask), out2[1] = sum(mask);
 * tensors [B,C,T], mask [B,T]; scratch = 512 floats.  Backward: any of the four gradient outputs may be NULL. */
int ttts_kl_loss(const float* z_p, const float* logs_q, const float* m_p, const float* logs_p, const float* mask, int32_t B, int32_t C,
                 int32_t T, float* scratch, float* out2, void* stream);
int ttts_kl_loss_bwd(const float* z_p, const float* m_p, const float* logs_p, const float* mask, const float* dL, const float* fwd_out2,
                     int32_t B, int32_t C, int32_t T, float* dz_p, float* dlogs_q, float* dm_p, float* dlogs_p, void* stream);

/* grouped Conv1d (DiscriminatorS, vq2.py:498-507: kernel 41, stride 4, four input channels per group; csrc/conv1d_grouped.cu, same validation
 * status): x [B,Cin,Tin], w [Cout, Cin/groups, K], dilation 1.  Backward: dx written, dw ACCUMULATES; either may be NULL. */
int ttts_gconv1d(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout, int32_t K,
                 int32_t stride, int32_t pad, int32_t groups, void* stream);
int ttts_gconv1d_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout,
                     int32_t K, int32_t stride, int32_t pad, int32_t groups, void* stream);
/* prior encoder enc_p_2 (TextEncoder + MRTE, vq2.py:17-164; csrc/text_encoder_kernels.cu, same validation status): multi-head attention over
 * short sequences, channel-major, with an optional windowed relative-position term (attentions.py:231-363 in closed form; emb_k / emb_v
 * [2 win + 1, C / heads] or NULL), query / key lengths and the reference's -1e4 masking; LayerNorm over the channel axis (modules.py:20-32). */
int ttts_attn_small(const float* q, const float* k, const float* v, const float* emb_k, const float* emb_v, const int64_t* q_len,
                    const int64_t* k_len, float* out, int32_t B, int32_t C, int32_t Tq, int32_t Tk, int32_t heads, int32_t win, void* stream);
int ttts_attn_small_bwd(const float* dout, const float* q, const float* k, const float* v, const float* emb_k, const float* emb_v,
                        const int64_t* q_len, const int64_t* k_len, float* dq, float* dk, float* dv, float* demb_k, float* demb_v,
                        int32_t B, int32_t C, int32_t Tq, int32_t Tk, int32_t heads, int32_t win, void* stream);   /* demb_* ACCUMULATE */
int ttts_layernorm_c(const float* x, const float* gamma, const float* beta, float* y, float* stats, int32_t B, int32_t C, int32_t T,
                     void* stream);                                                                                  /* stats: B*T*2 floats */
int ttts_layernorm_c_bwd(const float* dy, const float* x, const float* stats, const float* gamma, float* dx, float* dgamma, float* dbeta,
                         float* scratch, int32_t B, int32_t C, int32_t T, void* stream);                             /* scratch: B*T*2 floats */

/* --------------------------------------------------------------------------------------------
 * Diffusion mel-refiner train step (BASELINE config 5; csrc/diffusion_kernels.cu): the ops `AA_diffusion`
 * (ttts/diffusion/aa_model.py:69-287) and `SpacedDiffusion.training_losses` (ttts/utils/diffusion.py:903-1014) add to the training tape.
 * fp32, [B, C, T] channel-major.
 * ------------------------------------------------------------------------------------------ */
/* GroupNorm32 / `normalization` (ttts/utils/utils.py:119-137) fused with the ResBlock's scale-shift modulation and SiLU (aa_model.py:120-135):
 * y = act(GroupNorm_G(x; gamma, beta, eps 1e-5) * (1 + scale[b,c]) + shift[b,c]); scale / shift [B,C] or both NULL; act = SiLU if silu else
 * identity; stats [B,G,2] = (mean, rstd) out, read by the backward.  C / G <= 64. */
int ttts_groupnorm(const float* x, const float* gamma, const float* beta, const float* scale, const float* shift, float* y, float* stats,
                   int32_t B, int32_t C, int32_t T, int32_t G, int32_t silu, void* stream);
/* dx [B,C,T], dgamma / dbeta [C], dscale / dshift [B,C] (NULL without modulation) WRITTEN; scratch: B*C*2 floats */
int ttts_groupnorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* beta, const float* scale,
                       const float* shift, float* dx, float* dgamma, float* dbeta, float* dscale, float* dshift, float* scratch,
                       int32_t B, int32_t C, int32_t T, int32_t G, int32_t silu, void* stream);
/* nn.SiLU: out = x sigmoid(x) (backward = 0) or dy d/dx (backward = 1) */
int ttts_silu(const float* x, const float* dy, float* out, int64_t n, int32_t backward, void* stream);
/* QKVAttentionLegacy + RelativePositionBias (utils.py:148-175, utils/xtransformers.py:146-188) inside AttentionBlock: qkv [B,3C,T] with head h
 * owning channels [3 ch h, 3 ch (h+1)) = q | k | v (ch = C / H in {8,16,32,64}); scores = q.k / sqrt(ch) + sqrt(ch) table[diag[j-i+T-1], h],
 * non-causal softmax; out [B,C,T]; lse [B,H,T] out for the backward.  table [32,H]; diag int32 [2T-1] = T5 bucket of each relative position
 * (32 buckets, max_distance 64, bidirectional).  Flash-style: no [T,T] tensor is materialised. */
int ttts_attn_bias(const float* qkv, const float* table, const int32_t* diag, float* out, float* lse, int32_t B, int32_t C, int32_t T,
                   int32_t H, void* stream);
int64_t ttts_attn_bias_bwd_scratch_floats(int32_t B, int32_t T, int32_t H);
/* dqkv [B,3C,T] and dtable [32,H] WRITTEN (deterministic: per-CTA bucket sums reduced in fixed order) */
int ttts_attn_bias_bwd(const float* dout, const float* qkv, const float* out, const float* lse, const float* table, const int32_t* diag,
                       float* dqkv, float* dtable, float* scratch, int32_t B, int32_t C, int32_t T, int32_t H, void* stream);
/* GaussianDiffusion.q_sample (utils/diffusion.py:243-260): x_t = coef[b,0] x_start + coef[b,1] noise over [B, per] tensors.
 * coef [B,8] fp32 per sample = sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod,
 * posterior_mean_coef1, posterior_mean_coef2, posterior_log_variance_clipped, log(beta) at t[b] */
int ttts_diff_q_sample(const float* x_start, const float* noise, const float* coef, float* x_t, int32_t B, int64_t per, void* stream);
/* training_losses with epsilon prediction, learned-range variance, MSE loss (utils/diffusion.py:930-1014 with _vb_terms_bpd :903-928,
 * p_mean_variance :284-386, normal_kl / discretized_gaussian_log_likelihood :17-73): model_out [B, 2 Cn, T] = (eps | variance values),
 * x_start / x_t / noise [B,Cn,T], t_is0 int32 [B] -> terms [2,B] = (mse, vb), loss [1] = mean_b(mse + vb) (ttts/diffusion/train.py:172-180).
 * scratch: B * 64 floats.  Backward: d model_out for a gradient dL [1] of the loss (the mean prediction is detached inside vb). */
int ttts_diff_loss(const float* model_out, const float* x_start, const float* x_t, const float* noise, const float* coef, const int32_t* t_is0,
                   float* terms, float* loss, float* scratch, int32_t B, int32_t Cn, int32_t T, void* stream);
int ttts_diff_loss_bwd(const float* dL, const float* model_out, const float* x_start, const float* x_t, const float* noise, const float* coef,
                       const int32_t* t_is0, float* dout, int32_t B, int32_t Cn, int32_t T, void* stream);

/* Layout conversion around the tensor-core convolutions of the training tapes (the wide nn.Conv1d / Conv2d-(k,1) layers of aa_model.py:97-118,
 * 198-233 and vq2.py:341-416,418-496 as split-bf16 GEMMs on ttts_gemm_bf16): x [B,C,T] fp32 (optionally through leaky_relu 0.1) -> rows
 * [hi(x[b,:,t]) | lo(x[b,:,t])] (2C bf16) at row row_off + b rows_per_clip + t of a ZERO-INITIALISED buffer (the untouched rows are the zero
 * padding around and between the clips); and back: D fp32 (row row_off + b rows_per_clip + t, pitch ld) -> y [B,C,T].
 * dil > 1: sample t of clip b lives at row row_off + (b dil + t mod dil) rows_per_clip + t / dil instead -- de-interleaved, every residue class of
 * the time index a clip of its own, so that a convolution with dilation dil becomes an ordinary convolution over B dil clips. */
int ttts_cl_split(const float* x, void* out_bf16, int32_t B, int32_t C, int32_t T, int32_t rows_per_clip, int32_t row_off, int32_t lrelu, int32_t dil,
                  void* stream);
int ttts_cl_unpack(const float* D, float* y, int32_t B, int32_t C, int32_t T, int32_t ld, int32_t rows_per_clip, int32_t row_off,
                   const float* lrelu_x /* NULL, or [B,C,T]: y *= leaky_relu'(lrelu_x), slope 0.1 */, int32_t dil, void* stream);
/* Weight side of the same route: w [Cout,Cin,K] fp32 -> W1, W2 [R, K 2 Cr] bf16, per tap W1 = [hi | hi], W2 = [lo | 0] (the B operands of the GEMM
 * pair that reduces over all taps at once).  flip_transpose = 0: R = Cout, Cr = Cin; 1: the input-gradient form, R = Cin, Cr = Cout, taps reversed. */
int ttts_conv_w_concat(const float* w, void* W1, void* W2, int32_t Cout, int32_t Cin, int32_t K, int32_t flip_transpose, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTTS_B200_H */
