// Internal (C++) launch interface shared by the kernel translation units, the GPT engine and the C-ABI shims.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ttts_b200.h"

namespace ttts {

typedef __nv_bfloat16 bf16;

struct DropCfg {
    uint32_t thresh16;   // p * 65536 (0 = off)
    float scale;         // 1 / (1 - p)
    uint64_t seed;       // per (step, site, layer)
};
struct RowMap { int T, Tt, B; };   // T == 0: identity
inline RowMap no_map() { RowMap m; m.T = 0; m.Tt = 0; m.B = 0; return m; }
inline DropCfg no_drop() { DropCfg d; d.thresh16 = 0; d.scale = 1.f; d.seed = 0; return d; }

// gemm_tcgen05.cu
int gemm_bf16(const ttts_gemm_args& a, cudaStream_t stream);
int pick_split_k(int M, int N, int K);
bool use_2cta(int M, int N);
// gemm2_tcgen05.cu (CTA-pair kernel)
int gemm2_bf16(const ttts_gemm_args& a, bool pair, cudaStream_t stream);
int pick_split_k2(int M, int N, int K, bool pair);
bool use_legacy_gemm();

// elementwise.cu
int prep_tokens(const int64_t* text, int ld_text, int64_t* codes, int ld_codes, const int64_t* wav_lengths, int B, int TL, int CL,
                int mel_comp, int start_text, int stop_text, int start_mel, int stop_mel, int32_t* text_in, int32_t* text_tgt,
                int32_t* mel_in, int32_t* mel_tgt, cudaStream_t st);
int embed_fwd(const int32_t* text_in, const int32_t* mel_in, const float* Et, const float* Em, const float* Pt, const float* Pm,
              float* x, int B, int Tt, int Tm, int d, int Vt, int Vm, DropCfg drop, cudaStream_t st);
int embed_bwd(const int32_t* text_in, const int32_t* mel_in, const float* g, float* dEt, float* dEm, float* dPt, float* dPm, int B, int Tt,
              int Tm, int d, int Vt, int Vm, DropCfg drop, cudaStream_t st);
int ln_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, void* y, float* stats, int M, int d,
           bool dbl, bool out_bf16, RowMap map, cudaStream_t st);
int ln_bwd(const void* dy, int dy_is_f32, const float* x, const float* stats, const float* w1, const float* b1, const float* w2,
           const float* g_in, float* g_out, bf16* g16_out, float* dw1, float* db1, float* dw2, float* db2, float* dbias_next, int M, int d,
           bool dbl, DropCfg drop, RowMap map, cudaStream_t st);
int ce_fwd(const bf16* logits, int ld, int V, const int32_t* tgt, int rows, float* row_loss, float* row_lse, float* loss_out, cudaStream_t st);
int ce_bwd(const bf16* logits, int ld, int V, const int32_t* tgt, int rows, const float* row_lse, const float* gscale, float weight,
           bf16* dlogits, cudaStream_t st);
int colsum_bf16(const bf16* a, int ld, int M, int N, float* out, cudaStream_t st);
int cast_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st);
int grad_norm(const float* g, size_t n, float* partial, float* out_norm, cudaStream_t st);
int adamw_step(float* p, const float* g, float* m, float* v, bf16* p16, size_t n, const float* norm, float max_norm, float grad_scale,
               float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t st);

// attention.cu : qkv is [B*T, 3*d] bf16 (q | k | v, heads = contiguous 64-wide slices); o/do are [B*T, d] bf16.
int attn_fwd(const bf16* qkv, bf16* o, float* lse, int B, int T, int H, DropCfg drop, cudaStream_t st);
int attn_dropout_mask(uint8_t* mask, int BH, int T, DropCfg drop, cudaStream_t st);
int elem_dropout_mask(uint8_t* mask, int rows, int cols, DropCfg drop, cudaStream_t st);
int attn_bwd(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int B, int T, int H, DropCfg drop,
             cudaStream_t st);

// gpt_decode.cu : KV-cache decode step (inference_speech); gpt_layout.h owns the parameter layout, gpt_engine.cu the prefill
int64_t gpt_kv_bytes(int layers, int B, int H, int T_max);
int64_t gpt_decode_workspace_bytes(int B, int d);
int gpt_kv_fill_layer(const bf16* qkv, int B, int T, int d, int H, int n_pos, bf16* kcache, bf16* vcache, int T_max, cudaStream_t st);
int gpt_decode_step(const ttts_gpt_decode* a, cudaStream_t st);

}  // namespace ttts
