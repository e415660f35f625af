// Flat parameter layout of the UnifiedVoice GPT (reference state_dict tensors at 64-element-aligned offsets):
// shared by the train-step engine (gpt_engine.cu) and the KV-cache decode step (gpt_decode.cu).  Plain C++ (no CUDA), so the CPU
// emulation build of the decode kernels (tests/emu) uses the very same offsets.  Include after host_util.h (TTTS_CHECK_ARG).
#pragma once
#include <stdint.h>
#include "../../include/ttts_b200.h"

namespace ttts {

static inline int64_t pad64(int64_t n) { return (n + 63) / 64 * 64; }

struct ParamLayout {
    int64_t off[TTTS_P_COUNT];      // offset of tensor (layer 0 for per-layer tensors)
    int64_t numel[TTTS_P_COUNT];
    int64_t layer_stride, emb_end, top_begin, total;
};

static bool is_layer_tensor(int t) { return t >= TTTS_P_LN1_W && t <= TTTS_P_PR_B; }

static ParamLayout make_layout(const ttts_gpt_config& c) {
    ParamLayout L;
    const int64_t d = c.model_dim;
    int64_t o = 0;
    auto put = [&](int t, int64_t n) { L.off[t] = o; L.numel[t] = n; o += pad64(n); };
    put(TTTS_P_TEXT_EMB, (int64_t)c.n_text_vocab * d);
    put(TTTS_P_MEL_EMB, (int64_t)c.n_mel_vocab * d);
    put(TTTS_P_TEXT_POS, (int64_t)(c.max_text_tokens + 2) * d);
    put(TTTS_P_MEL_POS, (int64_t)(c.max_mel_tokens + 2) * d);
    L.emb_end = o;
    put(TTTS_P_LN1_W, d); put(TTTS_P_LN1_B, d);
    put(TTTS_P_ATTN_W, d * 3 * d); put(TTTS_P_ATTN_B, 3 * d);
    put(TTTS_P_PROJ_W, d * d); put(TTTS_P_PROJ_B, d);
    put(TTTS_P_LN2_W, d); put(TTTS_P_LN2_B, d);
    put(TTTS_P_FC_W, d * 4 * d); put(TTTS_P_FC_B, 4 * d);
    put(TTTS_P_PR_W, 4 * d * d); put(TTTS_P_PR_B, d);
    L.layer_stride = o - L.emb_end;
    o = L.emb_end + L.layer_stride * c.layers;
    L.top_begin = o;
    put(TTTS_P_LNF_W, d); put(TTTS_P_LNF_B, d); put(TTTS_P_FN_W, d); put(TTTS_P_FN_B, d);
    put(TTTS_P_TEXT_HEAD_W, (int64_t)c.n_text_vocab * d); put(TTTS_P_TEXT_HEAD_B, c.n_text_vocab);
    put(TTTS_P_MEL_HEAD_W, (int64_t)c.n_mel_vocab * d); put(TTTS_P_MEL_HEAD_B, c.n_mel_vocab);
    L.total = o;
    return L;
}
static inline int64_t poff(const ParamLayout& L, int t, int layer) { return L.off[t] + (is_layer_tensor(t) ? L.layer_stride * layer : 0); }

static int check_cfg(const ttts_gpt_config& c) {
    TTTS_CHECK_ARG(c.layers >= 1 && c.model_dim >= 128 && c.model_dim % 128 == 0 && c.model_dim <= 1024, "gpt: model_dim %d unsupported", c.model_dim);
    TTTS_CHECK_ARG(c.heads * 64 == c.model_dim, "gpt: only head_dim 64 is supported (heads=%d, model_dim=%d)", c.heads, c.model_dim);
    TTTS_CHECK_ARG(c.n_text_vocab > 1 && c.n_mel_vocab > 1, "gpt: bad vocab");
    return TTTS_OK;
}

// element offset of a tensor; -1 on a bad config / tensor / layer
static inline int64_t gpt_param_off(const ttts_gpt_config& c, int tensor, int layer) {
    if (tensor < 0 || tensor >= TTTS_P_COUNT || check_cfg(c) != TTTS_OK) return -1;
    if (is_layer_tensor(tensor) && (layer < 0 || layer >= c.layers)) return -1;
    return poff(make_layout(c), tensor, layer);
}

}  // namespace ttts
