// UnifiedVoice GPT train-step engine: the layer loop of ttts/gpt/model.py:453-510 (forward) and its autograd
// backward, sequenced as direct kernel launches on one stream -- no tracing, no recompute (activations for a
// 24L/d1024/B32/T1156 step are ~33 GB of the 180 GB HBM).
//
// Precision map: fp32 residual stream and LayerNorm statistics, bf16 tensor-core operands,
// fp32 TMEM accumulation, bf16 GEMM outputs (rounded exactly where autocast rounds), fp32 losses / grads / optimizer.
#include <string.h>
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#include "gpt_layout.h"

namespace ttts {

static inline int logits_ld(int V) { return (int)pad64(V); }

// ---------------------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------------------
struct Workspace {
    // sizes
    int B, TL, CL, Tt, Tm, T, M, d, L, H, Vt, Vm, ldt, ldm;
    bool save;
    int64_t total;
    // offsets (bytes)
    int64_t tok, resid, xmid, ln1s, ln2s, lnfs, h1, qkv, att, lse, h2, pre, act, enc, latent, logit_t, logit_m;
    int64_t rloss_t, rlse_t, rloss_m, rlse_m;
    int64_t g, g16, dpre, dqkv, datt, dhid, dlog_t, dlog_m, denc, delta;
    // per-layer strides (bytes) ; 0 when !save
    int64_t s_resid, s_xmid, s_ln, s_h, s_qkv, s_lse, s_big;
};

static Workspace carve(const ttts_gpt_config& c, int B, int TL, int CL, bool save) {
    Workspace w;
    memset(&w, 0, sizeof(w));
    w.B = B; w.TL = TL; w.CL = CL; w.Tt = TL + 2; w.Tm = CL + 2; w.T = w.Tt + w.Tm; w.M = B * w.T;
    w.d = c.model_dim; w.L = c.layers; w.H = c.heads; w.Vt = c.n_text_vocab; w.Vm = c.n_mel_vocab;
    w.ldt = logits_ld(w.Vt); w.ldm = logits_ld(w.Vm);
    w.save = save;
    const int64_t M = w.M, d = w.d;
    const int nl = save ? w.L : 1;
    int64_t o = 0;
    auto put = [&](int64_t bytes) { int64_t r = o; o += (bytes + 1023) / 1024 * 1024; return r; };
    w.tok = put((int64_t)2 * B * (w.Tt + w.Tm) * 4);
    // residual stream x_0..x_L (fp32).  Without saving we ping-pong two buffers.
    w.s_resid = M * d * 4;
    w.resid = put(w.s_resid * (save ? (w.L + 1) : 2));
    w.s_xmid = save ? M * d * 4 : 0;
    w.xmid = put(M * d * 4 * nl);
    w.s_ln = save ? M * 2 * 4 : 0;
    w.ln1s = put(M * 2 * 4 * nl);
    w.ln2s = put(M * 2 * 4 * nl);
    w.lnfs = put(M * 4 * 4);
    w.s_h = save ? M * d * 2 : 0;
    w.h1 = put(M * d * 2 * nl);
    w.h2 = put(M * d * 2 * nl);
    w.att = put(M * d * 2 * nl);
    w.s_qkv = save ? M * 3 * d * 2 : 0;
    w.qkv = put(M * 3 * d * 2 * nl);
    w.s_lse = save ? (int64_t)B * w.H * w.T * 4 : 0;
    w.lse = put((int64_t)B * w.H * w.T * 4 * nl);
    w.s_big = save ? M * 4 * d * 2 : 0;
    w.pre = put(M * 4 * d * 2 * nl);
    w.act = put(M * 4 * d * 2 * nl);
    w.enc = put(M * d * 2);
    w.latent = put(M * d * 4);
    w.logit_t = put((int64_t)B * w.Tt * w.ldt * 2);
    w.logit_m = put((int64_t)B * w.Tm * w.ldm * 2);
    w.rloss_t = put((int64_t)B * w.Tt * 4); w.rlse_t = put((int64_t)B * w.Tt * 4);
    w.rloss_m = put((int64_t)B * w.Tm * 4); w.rlse_m = put((int64_t)B * w.Tm * 4);
    if (save) {
        w.g = put(M * d * 4);
        w.g16 = put(M * d * 2);
        w.dpre = put(M * 4 * d * 2);
        w.dqkv = put(M * 3 * d * 2);
        w.datt = put(M * d * 2);
        w.dhid = put(M * d * 2);
        w.dlog_t = put((int64_t)B * w.Tt * w.ldt * 2);
        w.dlog_m = put((int64_t)B * w.Tm * w.ldm * 2);
        w.denc = put(M * d * 2);
        w.delta = put(((int64_t)B * w.H * w.T + 64 + M * d) * 4);     // attention bwd scratch: delta | fp32 dQ accumulator
    }
    w.total = o;
    return w;
}

static inline DropCfg site_drop(float p, uint64_t seed, int site, int layer) {
    DropCfg dc = no_drop();
    if (p > 0.f) {
        dc.thresh16 = 2u * (uint32_t)(p * 32768.0f + 0.5f);         // p as a multiple of 2^-15 (the attention hash decides on 15-bit fields)
        dc.scale = 1.0f / (1.0f - (float)dc.thresh16 / 65536.0f);
        uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(site * 1024 + layer + 1);
        z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL; z ^= z >> 27; z *= 0x94d049bb133111ebULL; z ^= z >> 31;
        dc.seed = z;
    }
    return dc;
}
enum { SITE_EMBD = 0, SITE_ATTN_P = 1, SITE_ATTN_O = 2, SITE_MLP_O = 3 };

static int gemm(int M, int N, int K, const void* A, int lda, bool a_mn, const void* B, int ldb, bool b_mn, int epi, void* out, int ldo,
                const float* bias, const void* aux, int ldaux, void* aux_out, int ldaux_out, int split_k, DropCfg drop, cudaStream_t st) {
    ttts_gemm_args g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.a_mn = a_mn; g.B = B; g.ldb = ldb; g.b_mn = b_mn;
    g.epi = epi; g.out = out; g.ldo = ldo; g.bias = bias; g.aux = aux; g.ldaux = ldaux; g.aux_out = aux_out; g.ldaux_out = ldaux_out;
    g.split_k = split_k;
    g.drop_thresh16 = drop.thresh16; g.drop_scale = drop.scale; g.drop_seed = drop.seed;
    return gemm_bf16(g, st);
}
// weight gradient: dW[Kin, Nout] += X[tokens, Kin]^T * dY[tokens, Nout]
static int wgrad(int Kin, int Nout, int tokens, const bf16* X, int ldx, const bf16* dY, int ldy, float* dW, int ldw, cudaStream_t st) {
    return gemm(Kin, Nout, tokens, X, ldx, true, dY, ldy, true, TTTS_EPI_F32_ADD, dW, ldw, nullptr, nullptr, 0, nullptr, 0,
                pick_split_k(Kin, Nout, tokens), no_drop(), st);
}

static int validate_io(const ttts_gpt_io* io, Workspace& w, ParamLayout& P) {
    TTTS_CHECK_ARG(io != nullptr, "gpt: null io");
    TTTS_RUN(check_cfg(io->cfg));
    TTTS_CHECK_ARG(io->B >= 1 && io->TL >= 0 && io->CL >= 0, "gpt: bad batch shape");
    TTTS_CHECK_ARG(io->TL <= io->cfg.max_text_tokens && io->CL <= io->cfg.max_mel_tokens,
                   "gpt: sequence (%d text, %d codes) exceeds the position tables (%d, %d)", io->TL, io->CL, io->cfg.max_text_tokens,
                   io->cfg.max_mel_tokens);
    TTTS_CHECK_ARG(io->params && io->params16 && io->workspace, "gpt: null buffer");
    w = carve(io->cfg, io->B, io->TL, io->CL, io->save_acts != 0);
    TTTS_CHECK_ARG(io->workspace_bytes >= w.total, "gpt: workspace too small (%lld < %lld)", (long long)io->workspace_bytes, (long long)w.total);
    TTTS_CHECK_ARG(((uintptr_t)io->workspace & 255) == 0, "gpt: workspace must be 256B aligned");
    P = make_layout(io->cfg);
    return TTTS_OK;
}

int gpt_forward(const ttts_gpt_io* io, cudaStream_t st) {
    Workspace w; ParamLayout P;
    TTTS_RUN(validate_io(io, w, P));
    const ttts_gpt_config& c = io->cfg;
    uint8_t* ws = reinterpret_cast<uint8_t*>(io->workspace);
    const float* p32 = io->params;
    const bf16* p16 = reinterpret_cast<const bf16*>(io->params16);
    const int M = w.M, d = w.d, B = w.B;
    const float dp = io->drop_p;

    int32_t* text_in = reinterpret_cast<int32_t*>(ws + w.tok);
    int32_t* text_tgt = text_in + B * w.Tt;
    int32_t* mel_in = text_tgt + B * w.Tt;
    int32_t* mel_tgt = mel_in + B * w.Tm;
    TTTS_CHECK_ARG(io->text && io->codes && io->wav_lengths, "gpt: null token inputs");
    TTTS_RUN(prep_tokens(io->text, io->ld_text, io->codes, io->ld_codes, io->wav_lengths, B, w.TL, w.CL, c.mel_length_compression,
                         c.start_text_token, c.stop_text_token, c.start_mel_token, c.stop_mel_token, text_in, text_tgt, mel_in, mel_tgt, st));

    auto resid = [&](int l) { return reinterpret_cast<float*>(ws + w.resid + w.s_resid * (w.save ? l : (l & 1))); };
    NvtxRange nv_fwd("ttts.gpt.forward");
    TTTS_RUN(embed_fwd(text_in, mel_in, p32 + P.off[TTTS_P_TEXT_EMB], p32 + P.off[TTTS_P_MEL_EMB], p32 + P.off[TTTS_P_TEXT_POS],
                       p32 + P.off[TTTS_P_MEL_POS], resid(0), B, w.Tt, w.Tm, d, w.Vt, w.Vm, site_drop(dp, io->seed, SITE_EMBD, 0), st));

    for (int l = 0; l < w.L; ++l) {
        NvtxRange nv_layer("ttts.gpt.fwd.layer", l);
        float* x = resid(l);
        float* xmid = reinterpret_cast<float*>(ws + w.xmid + w.s_xmid * l);
        float* xnext = resid(l + 1);
        bf16* h1 = reinterpret_cast<bf16*>(ws + w.h1 + w.s_h * l);
        bf16* h2 = reinterpret_cast<bf16*>(ws + w.h2 + w.s_h * l);
        bf16* att = reinterpret_cast<bf16*>(ws + w.att + w.s_h * l);
        bf16* qkv = reinterpret_cast<bf16*>(ws + w.qkv + w.s_qkv * l);
        float* lse = reinterpret_cast<float*>(ws + w.lse + w.s_lse * l);
        bf16* pre = reinterpret_cast<bf16*>(ws + w.pre + w.s_big * l);
        bf16* act = reinterpret_cast<bf16*>(ws + w.act + w.s_big * l);
        float* ln1s = reinterpret_cast<float*>(ws + w.ln1s + w.s_ln * l);
        float* ln2s = reinterpret_cast<float*>(ws + w.ln2s + w.s_ln * l);

        // ln_1 -> c_attn (+bias)                                  HF: modeling_gpt2.py:273,185
        TTTS_RUN(ln_fwd(x, p32 + poff(P, TTTS_P_LN1_W, l), p32 + poff(P, TTTS_P_LN1_B, l), nullptr, nullptr, h1, ln1s, M, d, false, true, no_map(), st));
        TTTS_RUN(gemm(M, 3 * d, d, h1, d, false, p16 + poff(P, TTTS_P_ATTN_W, l), 3 * d, true, TTTS_EPI_BF16, qkv, 3 * d,
                      p32 + poff(P, TTTS_P_ATTN_B, l), nullptr, 0, nullptr, 0, 1, no_drop(), st));
        // causal attention (+ attn dropout)                       HF: modeling_gpt2.py:185-220
        TTTS_RUN(attn_fwd(qkv, att, lse, B, w.T, w.H, site_drop(dp, io->seed, SITE_ATTN_P, l), st));
        // c_proj + resid dropout + residual                       HF: modeling_gpt2.py:223-224,282
        TTTS_RUN(gemm(M, d, d, att, d, false, p16 + poff(P, TTTS_P_PROJ_W, l), d, true, TTTS_EPI_RESID, xmid, d,
                      p32 + poff(P, TTTS_P_PROJ_B, l), x, d, nullptr, 0, 1, site_drop(dp, io->seed, SITE_ATTN_O, l), st));
        // ln_2 -> c_fc -> gelu_new                                HF: modeling_gpt2.py:304-305,239-240
        TTTS_RUN(ln_fwd(xmid, p32 + poff(P, TTTS_P_LN2_W, l), p32 + poff(P, TTTS_P_LN2_B, l), nullptr, nullptr, h2, ln2s, M, d, false, true, no_map(), st));
        TTTS_RUN(gemm(M, 4 * d, d, h2, d, false, p16 + poff(P, TTTS_P_FC_W, l), 4 * d, true, TTTS_EPI_GELU, act, 4 * d,
                      p32 + poff(P, TTTS_P_FC_B, l), nullptr, 0, pre, 4 * d, 1, no_drop(), st));
        // mlp c_proj + dropout + residual                         HF: modeling_gpt2.py:241-242,307
        TTTS_RUN(gemm(M, d, 4 * d, act, 4 * d, false, p16 + poff(P, TTTS_P_PR_W, l), d, true, TTTS_EPI_RESID, xnext, d,
                      p32 + poff(P, TTTS_P_PR_B, l), xmid, d, nullptr, 0, 1, site_drop(dp, io->seed, SITE_MLP_O, l), st));
    }

    float* xL = resid(w.L);
    float* lnfs = reinterpret_cast<float*>(ws + w.lnfs);
    if (io->want_latent) {
        // return_latent path: fp32 final_norm(ln_f(x)) in [b, t] row order        ttts/gpt/model.py:426-430
        TTTS_RUN(ln_fwd(xL, p32 + P.off[TTTS_P_LNF_W], p32 + P.off[TTTS_P_LNF_B], p32 + P.off[TTTS_P_FN_W], p32 + P.off[TTTS_P_FN_B],
                        ws + w.latent, lnfs, M, d, true, false, no_map(), st));
        return TTTS_OK;
    }
    // ln_f -> final_norm (one fused kernel), rows regrouped [text rows ; mel rows] for the two heads
    RowMap map; map.T = w.T; map.Tt = w.Tt; map.B = B;
    bf16* enc = reinterpret_cast<bf16*>(ws + w.enc);
    TTTS_RUN(ln_fwd(xL, p32 + P.off[TTTS_P_LNF_W], p32 + P.off[TTTS_P_LNF_B], p32 + P.off[TTTS_P_FN_W], p32 + P.off[TTTS_P_FN_B], enc, lnfs,
                    M, d, true, true, map, st));
    bf16* enc_t = enc;
    bf16* enc_m = enc + (size_t)B * w.Tt * d;
    bf16* logit_t = reinterpret_cast<bf16*>(ws + w.logit_t);
    bf16* logit_m = reinterpret_cast<bf16*>(ws + w.logit_m);
    // heads (nn.Linear layout [V, d] = K-major B)                 ttts/gpt/model.py:432-437
    TTTS_RUN(gemm(B * w.Tt, w.Vt, d, enc_t, d, false, p16 + P.off[TTTS_P_TEXT_HEAD_W], d, false, TTTS_EPI_BF16, logit_t, w.ldt,
                  p32 + P.off[TTTS_P_TEXT_HEAD_B], nullptr, 0, nullptr, 0, 1, no_drop(), st));
    TTTS_RUN(gemm(B * w.Tm, w.Vm, d, enc_m, d, false, p16 + P.off[TTTS_P_MEL_HEAD_W], d, false, TTTS_EPI_BF16, logit_m, w.ldm,
                  p32 + P.off[TTTS_P_MEL_HEAD_B], nullptr, 0, nullptr, 0, 1, no_drop(), st));
    // mean cross-entropy over ALL positions                       ttts/gpt/model.py:508-509
    TTTS_CHECK_ARG(io->losses != nullptr, "gpt: null losses");
    TTTS_RUN(ce_fwd(logit_t, w.ldt, w.Vt, text_tgt, B * w.Tt, reinterpret_cast<float*>(ws + w.rloss_t), reinterpret_cast<float*>(ws + w.rlse_t),
                    io->losses, st));
    TTTS_RUN(ce_fwd(logit_m, w.ldm, w.Vm, mel_tgt, B * w.Tm, reinterpret_cast<float*>(ws + w.rloss_m), reinterpret_cast<float*>(ws + w.rlse_m),
                    io->losses + 1, st));
    return TTTS_OK;
}

int gpt_backward(const ttts_gpt_io* io, int stage_begin, int stage_end, cudaStream_t st) {
    Workspace w; ParamLayout P;
    TTTS_RUN(validate_io(io, w, P));
    TTTS_CHECK_ARG(io->save_acts, "gpt backward needs a forward run with save_acts=1");
    TTTS_CHECK_ARG(io->grads != nullptr, "gpt backward: null grads");
    TTTS_CHECK_ARG(stage_begin >= 0 && stage_end <= w.L + 2 && stage_begin <= stage_end, "gpt backward: bad stage range");
    uint8_t* ws = reinterpret_cast<uint8_t*>(io->workspace);
    const float* p32 = io->params;
    const bf16* p16 = reinterpret_cast<const bf16*>(io->params16);
    float* gr = io->grads;
    const int M = w.M, d = w.d, B = w.B;
    const float dp = io->drop_p;

    int32_t* text_in = reinterpret_cast<int32_t*>(ws + w.tok);
    int32_t* text_tgt = text_in + B * w.Tt;
    int32_t* mel_in = text_tgt + B * w.Tt;
    int32_t* mel_tgt = mel_in + B * w.Tm;
    float* g = reinterpret_cast<float*>(ws + w.g);
    bf16* g16 = reinterpret_cast<bf16*>(ws + w.g16);
    bf16* dpre = reinterpret_cast<bf16*>(ws + w.dpre);
    bf16* dqkv = reinterpret_cast<bf16*>(ws + w.dqkv);
    bf16* datt = reinterpret_cast<bf16*>(ws + w.datt);
    bf16* dhid = reinterpret_cast<bf16*>(ws + w.dhid);
    float* delta = reinterpret_cast<float*>(ws + w.delta);

    NvtxRange nv_bwd("ttts.gpt.backward");
    for (int stage = stage_begin; stage < stage_end; ++stage) {
        NvtxRange nv_stage(stage == 0 ? "ttts.gpt.bwd.heads" : (stage <= w.L ? "ttts.gpt.bwd.layer" : "ttts.gpt.bwd.embeddings"), stage <= w.L ? w.L - stage : 0);
        if (stage == 0) {
            // ---------------- heads + CE + final double LayerNorm ----------------
            bf16* enc = reinterpret_cast<bf16*>(ws + w.enc);
            bf16* enc_t = enc;
            bf16* enc_m = enc + (size_t)B * w.Tt * d;
            bf16* logit_t = reinterpret_cast<bf16*>(ws + w.logit_t);
            bf16* logit_m = reinterpret_cast<bf16*>(ws + w.logit_m);
            bf16* dlog_t = reinterpret_cast<bf16*>(ws + w.dlog_t);
            bf16* dlog_m = reinterpret_cast<bf16*>(ws + w.dlog_m);
            bf16* denc = reinterpret_cast<bf16*>(ws + w.denc);
            bf16* denc_t = denc;
            bf16* denc_m = denc + (size_t)B * w.Tt * d;
            TTTS_RUN(ce_bwd(logit_t, w.ldt, w.Vt, text_tgt, B * w.Tt, reinterpret_cast<float*>(ws + w.rlse_t), io->gscale_text, io->weight_text, dlog_t, st));
            TTTS_RUN(ce_bwd(logit_m, w.ldm, w.Vm, mel_tgt, B * w.Tm, reinterpret_cast<float*>(ws + w.rlse_m), io->gscale_mel, io->weight_mel, dlog_m, st));
            // dW_head[V, d] += dlogits^T enc ; db += colsum(dlogits) ; denc = dlogits W_head
            TTTS_RUN(wgrad(w.Vt, d, B * w.Tt, dlog_t, w.ldt, enc_t, d, gr + P.off[TTTS_P_TEXT_HEAD_W], d, st));
            TTTS_RUN(wgrad(w.Vm, d, B * w.Tm, dlog_m, w.ldm, enc_m, d, gr + P.off[TTTS_P_MEL_HEAD_W], d, st));
            TTTS_RUN(colsum_bf16(dlog_t, w.ldt, B * w.Tt, w.Vt, gr + P.off[TTTS_P_TEXT_HEAD_B], st));
            TTTS_RUN(colsum_bf16(dlog_m, w.ldm, B * w.Tm, w.Vm, gr + P.off[TTTS_P_MEL_HEAD_B], st));
            TTTS_RUN(gemm(B * w.Tt, d, w.Vt, dlog_t, w.ldt, false, p16 + P.off[TTTS_P_TEXT_HEAD_W], d, true, TTTS_EPI_BF16, denc_t, d, nullptr,
                          nullptr, 0, nullptr, 0, 1, no_drop(), st));
            TTTS_RUN(gemm(B * w.Tm, d, w.Vm, dlog_m, w.ldm, false, p16 + P.off[TTTS_P_MEL_HEAD_W], d, true, TTTS_EPI_BF16, denc_m, d, nullptr,
                          nullptr, 0, nullptr, 0, 1, no_drop(), st));
            RowMap map; map.T = w.T; map.Tt = w.Tt; map.B = B;
            float* xL = reinterpret_cast<float*>(ws + w.resid + w.s_resid * w.L);
            TTTS_RUN(ln_bwd(denc, 0, xL, reinterpret_cast<float*>(ws + w.lnfs), p32 + P.off[TTTS_P_LNF_W], p32 + P.off[TTTS_P_LNF_B],
                            p32 + P.off[TTTS_P_FN_W], nullptr, g, g16, gr + P.off[TTTS_P_LNF_W], gr + P.off[TTTS_P_LNF_B], gr + P.off[TTTS_P_FN_W],
                            gr + P.off[TTTS_P_FN_B], gr + poff(P, TTTS_P_PR_B, w.L - 1), M, d, true, site_drop(dp, io->seed, SITE_MLP_O, w.L - 1),
                            map, st));
        } else if (stage <= w.L) {
            const int l = w.L - stage;
            float* x = reinterpret_cast<float*>(ws + w.resid + w.s_resid * l);
            float* xmid = reinterpret_cast<float*>(ws + w.xmid + w.s_xmid * l);
            bf16* h1 = reinterpret_cast<bf16*>(ws + w.h1 + w.s_h * l);
            bf16* h2 = reinterpret_cast<bf16*>(ws + w.h2 + w.s_h * l);
            bf16* att = reinterpret_cast<bf16*>(ws + w.att + w.s_h * l);
            bf16* qkv = reinterpret_cast<bf16*>(ws + w.qkv + w.s_qkv * l);
            float* lse = reinterpret_cast<float*>(ws + w.lse + w.s_lse * l);
            bf16* pre = reinterpret_cast<bf16*>(ws + w.pre + w.s_big * l);
            bf16* act = reinterpret_cast<bf16*>(ws + w.act + w.s_big * l);
            float* ln1s = reinterpret_cast<float*>(ws + w.ln1s + w.s_ln * l);
            float* ln2s = reinterpret_cast<float*>(ws + w.ln2s + w.s_ln * l);
            // g16 currently holds bf16(dropmask_mlp(g)) = gradient of the mlp c_proj output
            // ---- mlp c_proj ----
            TTTS_RUN(wgrad(4 * d, d, M, act, 4 * d, g16, d, gr + poff(P, TTTS_P_PR_W, l), d, st));
            TTTS_RUN(gemm(M, 4 * d, d, g16, d, false, p16 + poff(P, TTTS_P_PR_W, l), d, false, TTTS_EPI_DGELU, dpre, 4 * d, nullptr, pre, 4 * d,
                          nullptr, 0, 1, no_drop(), st));
            // ---- c_fc ----
            TTTS_RUN(colsum_bf16(dpre, 4 * d, M, 4 * d, gr + poff(P, TTTS_P_FC_B, l), st));
            TTTS_RUN(wgrad(d, 4 * d, M, h2, d, dpre, 4 * d, gr + poff(P, TTTS_P_FC_W, l), 4 * d, st));
            TTTS_RUN(gemm(M, d, 4 * d, dpre, 4 * d, false, p16 + poff(P, TTTS_P_FC_W, l), 4 * d, false, TTTS_EPI_BF16, dhid, d, nullptr, nullptr, 0,
                          nullptr, 0, 1, no_drop(), st));
            // ---- ln_2 ---- g += LNbwd(dhid) ; g16 = bf16(dropmask_attn_out(g)) ; db(attn c_proj) = colsum(g16)
            TTTS_RUN(ln_bwd(dhid, 0, xmid, ln2s, p32 + poff(P, TTTS_P_LN2_W, l), p32 + poff(P, TTTS_P_LN2_B, l), nullptr, g, g, g16,
                            gr + poff(P, TTTS_P_LN2_W, l), gr + poff(P, TTTS_P_LN2_B, l), nullptr, nullptr, gr + poff(P, TTTS_P_PROJ_B, l), M, d, false,
                            site_drop(dp, io->seed, SITE_ATTN_O, l), no_map(), st));
            // ---- attn c_proj ----
            TTTS_RUN(wgrad(d, d, M, att, d, g16, d, gr + poff(P, TTTS_P_PROJ_W, l), d, st));
            TTTS_RUN(gemm(M, d, d, g16, d, false, p16 + poff(P, TTTS_P_PROJ_W, l), d, false, TTTS_EPI_BF16, datt, d, nullptr, nullptr, 0, nullptr, 0, 1,
                          no_drop(), st));
            // ---- attention ----
            TTTS_RUN(attn_bwd(qkv, att, datt, lse, delta, dqkv, B, w.T, w.H, site_drop(dp, io->seed, SITE_ATTN_P, l), st));
            // ---- c_attn ----
            TTTS_RUN(colsum_bf16(dqkv, 3 * d, M, 3 * d, gr + poff(P, TTTS_P_ATTN_B, l), st));
            TTTS_RUN(wgrad(d, 3 * d, M, h1, d, dqkv, 3 * d, gr + poff(P, TTTS_P_ATTN_W, l), 3 * d, st));
            TTTS_RUN(gemm(M, d, 3 * d, dqkv, 3 * d, false, p16 + poff(P, TTTS_P_ATTN_W, l), 3 * d, false, TTTS_EPI_BF16, dhid, d, nullptr, nullptr, 0,
                          nullptr, 0, 1, no_drop(), st));
            // ---- ln_1 ---- g += LNbwd(dhid) ; for l > 0 also g16 / db for the previous layer's mlp c_proj
            TTTS_RUN(ln_bwd(dhid, 0, x, ln1s, p32 + poff(P, TTTS_P_LN1_W, l), p32 + poff(P, TTTS_P_LN1_B, l), nullptr, g, g, l > 0 ? g16 : nullptr,
                            gr + poff(P, TTTS_P_LN1_W, l), gr + poff(P, TTTS_P_LN1_B, l), nullptr, nullptr, l > 0 ? gr + poff(P, TTTS_P_PR_B, l - 1) : nullptr,
                            M, d, false, l > 0 ? site_drop(dp, io->seed, SITE_MLP_O, l - 1) : no_drop(), no_map(), st));
        } else {
            // ---------------- embeddings ----------------
            TTTS_RUN(embed_bwd(text_in, mel_in, g, gr + P.off[TTTS_P_TEXT_EMB], gr + P.off[TTTS_P_MEL_EMB], gr + P.off[TTTS_P_TEXT_POS],
                               gr + P.off[TTTS_P_MEL_POS], B, w.Tt, w.Tm, d, w.Vt, w.Vm, site_drop(dp, io->seed, SITE_EMBD, 0), st));
        }
    }
    return TTTS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// KV-cache decode support (kernels in gpt_decode.cu): the cache fill from a saved forward
// ---------------------------------------------------------------------------------------------------------
int gpt_kv_prefill(const ttts_gpt_io* io, void* kv, int64_t kv_bytes, int T_max, int n_pos, cudaStream_t st) {
    Workspace w; ParamLayout P;
    TTTS_RUN(validate_io(io, w, P));
    TTTS_CHECK_ARG(io->save_acts, "kv prefill needs a forward run with save_acts=1 (every layer's c_attn output is kept)");
    TTTS_CHECK_ARG(kv != nullptr && ((uintptr_t)kv & 15) == 0, "kv prefill: null / unaligned cache");
    TTTS_CHECK_ARG(T_max >= n_pos && n_pos >= 1 && n_pos <= w.T, "kv prefill: %d positions (sequence %d, cache capacity %d)", n_pos, w.T, T_max);
    TTTS_CHECK_ARG(kv_bytes >= gpt_kv_bytes(w.L, w.B, w.H, T_max), "kv prefill: cache too small (%lld < %lld)", (long long)kv_bytes,
                   (long long)gpt_kv_bytes(w.L, w.B, w.H, T_max));
    const uint8_t* ws = reinterpret_cast<const uint8_t*>(io->workspace);
    bf16* cache = reinterpret_cast<bf16*>(kv);
    const size_t half = (size_t)w.B * w.H * T_max * 64;
    for (int l = 0; l < w.L; ++l) {
        const bf16* qkv = reinterpret_cast<const bf16*>(ws + w.qkv + w.s_qkv * l);
        TTTS_RUN(gpt_kv_fill_layer(qkv, w.B, w.T, w.d, w.H, n_pos, cache + (size_t)l * 2 * half, cache + (size_t)l * 2 * half + half, T_max, st));
    }
    return TTTS_OK;
}

}  // namespace ttts

using namespace ttts;

extern "C" {

int64_t ttts_gpt_param_offset(const ttts_gpt_config* cfg, int32_t tensor, int32_t layer) {
    if (!cfg || tensor < 0 || tensor >= TTTS_P_COUNT || check_cfg(*cfg) != TTTS_OK) return -1;
    if (is_layer_tensor(tensor) && (layer < 0 || layer >= cfg->layers)) return -1;
    ParamLayout L = make_layout(*cfg);
    return poff(L, tensor, layer);
}
int64_t ttts_gpt_param_numel(const ttts_gpt_config* cfg, int32_t tensor) {
    if (!cfg || tensor < 0 || tensor >= TTTS_P_COUNT || check_cfg(*cfg) != TTTS_OK) return -1;
    return make_layout(*cfg).numel[tensor];
}
int64_t ttts_gpt_param_count(const ttts_gpt_config* cfg) {
    if (!cfg || check_cfg(*cfg) != TTTS_OK) return -1;
    return make_layout(*cfg).total;
}
int32_t ttts_gpt_stage_range(const ttts_gpt_config* cfg, int32_t stage, int64_t* begin, int64_t* end) {
    if (!cfg || !begin || !end || check_cfg(*cfg) != TTTS_OK) return TTTS_ERR_INVALID;
    ParamLayout L = make_layout(*cfg);
    if (stage == 0) { *begin = L.top_begin; *end = L.total; }
    else if (stage <= cfg->layers) { int l = cfg->layers - stage; *begin = L.emb_end + L.layer_stride * l; *end = *begin + L.layer_stride; }
    else if (stage == cfg->layers + 1) { *begin = 0; *end = L.emb_end; }
    else { set_error("gpt: bad stage %d", stage); return TTTS_ERR_INVALID; }
    return TTTS_OK;
}
int64_t ttts_gpt_workspace_bytes(const ttts_gpt_config* cfg, int32_t B, int32_t TL, int32_t CL, int32_t save_acts) {
    if (!cfg || check_cfg(*cfg) != TTTS_OK || B < 1 || TL < 0 || CL < 0) return -1;
    return carve(*cfg, B, TL, CL, save_acts != 0).total;
}
int32_t ttts_gpt_logits_ld(int32_t vocab) { return logits_ld(vocab); }
int64_t ttts_gpt_workspace_offset(const ttts_gpt_config* cfg, int32_t B, int32_t TL, int32_t CL, int32_t save_acts, int32_t item, int32_t layer) {
    if (!cfg || check_cfg(*cfg) != TTTS_OK || B < 1) return -1;
    Workspace w = carve(*cfg, B, TL, CL, save_acts != 0);
    switch (item) {
    case TTTS_WS_MEL_LOGITS: return w.logit_m;
    case TTTS_WS_TEXT_LOGITS: return w.logit_t;
    case TTTS_WS_LATENT: return w.latent;
    case TTTS_WS_RESID:
        if (layer < 0 || layer > cfg->layers) return -1;
        return w.resid + w.s_resid * (w.save ? layer : (layer & 1));
    case TTTS_WS_TOKENS: return w.tok;
    default: return -1;
    }
}

int ttts_gpt_forward(const ttts_gpt_io* io, void* stream) { return gpt_forward(io, (cudaStream_t)stream); }
int ttts_gpt_backward(const ttts_gpt_io* io, int32_t stage_begin, int32_t stage_end, void* stream) {
    return gpt_backward(io, stage_begin, stage_end, (cudaStream_t)stream);
}

int64_t ttts_gpt_kv_bytes(const ttts_gpt_config* cfg, int32_t B, int32_t T_max) {
    if (!cfg || check_cfg(*cfg) != TTTS_OK || B < 1 || T_max < 1) return -1;
    return gpt_kv_bytes(cfg->layers, B, cfg->heads, T_max);
}
int64_t ttts_gpt_decode_workspace_bytes(const ttts_gpt_config* cfg, int32_t B) {
    if (!cfg || check_cfg(*cfg) != TTTS_OK || B < 1) return -1;
    return gpt_decode_workspace_bytes(B, cfg->model_dim);
}
int ttts_gpt_kv_prefill(const ttts_gpt_io* io, void* kv, int64_t kv_bytes, int32_t T_max, int32_t n_pos, void* stream) {
    return gpt_kv_prefill(io, kv, kv_bytes, T_max, n_pos, (cudaStream_t)stream);
}
int ttts_gpt_decode_step(const ttts_gpt_decode* args, void* stream) { return gpt_decode_step(args, (cudaStream_t)stream); }

int ttts_cast_bf16(const float* src, void* dst, int64_t n, void* stream) { return cast_bf16(src, (bf16*)dst, (size_t)n, (cudaStream_t)stream); }
int ttts_grad_norm(const float* grads, int64_t n, float* scratch, float* norm_out, void* stream) {
    return grad_norm(grads, (size_t)n, scratch, norm_out, (cudaStream_t)stream);
}
int ttts_adamw_step(float* params, const float* grads, float* m, float* v, void* p16, int64_t n, const float* norm, float max_norm,
                    float grad_scale, float lr, float beta1, float beta2, float eps, float wd, int32_t step, void* stream) {
    return adamw_step(params, grads, m, v, (bf16*)p16, (size_t)n, norm, max_norm, grad_scale, lr, beta1, beta2, eps, wd, step, (cudaStream_t)stream);
}

int ttts_layernorm_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, void* y, float* stats, int32_t M,
                       int32_t d, int32_t dbl, int32_t out_bf16, void* stream) {
    return ln_fwd(x, w1, b1, w2, b2, y, stats, M, d, dbl != 0, out_bf16 != 0, no_map(), (cudaStream_t)stream);
}
int ttts_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, const float* stats, const float* w1, const float* b1, const float* w2,
                       const float* g_in, float* g_out, void* g16_out, float* dw1, float* db1, float* dw2, float* db2, float* dbias_next,
                       int32_t M, int32_t d, int32_t dbl, void* stream) {
    return ln_bwd(dy, dy_is_f32, x, stats, w1, b1, w2, g_in, g_out, (bf16*)g16_out, dw1, db1, dw2, db2, dbias_next, M, d, dbl != 0, no_drop(),
                  no_map(), (cudaStream_t)stream);
}
static DropCfg user_drop(float p, uint64_t seed) {
    DropCfg dc = no_drop();
    if (p > 0.f) { dc.thresh16 = 2u * (uint32_t)(p * 32768.0f + 0.5f); dc.scale = 1.0f / (1.0f - (float)dc.thresh16 / 65536.0f); dc.seed = seed; }
    return dc;
}
int64_t ttts_attn_bwd_scratch_floats(int32_t B, int32_t T, int32_t H) { return (int64_t)B * H * T + 64 + (int64_t)B * T * H * 64; }
int ttts_attn_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t T, int32_t H, float drop_p, uint64_t seed, void* stream) {
    return attn_fwd((const bf16*)qkv, (bf16*)out, lse, B, T, H, user_drop(drop_p, seed), (cudaStream_t)stream);
}
int ttts_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int32_t B, int32_t T, int32_t H,
                  float drop_p, uint64_t seed, void* stream) {
    return attn_bwd((const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, delta, (bf16*)dqkv, B, T, H, user_drop(drop_p, seed),
                    (cudaStream_t)stream);
}
int ttts_attn_dropout_mask(uint8_t* mask, int32_t BH, int32_t T, float drop_p, uint64_t seed, void* stream) {
    return attn_dropout_mask(mask, BH, T, user_drop(drop_p, seed), (cudaStream_t)stream);
}
int ttts_gpt_dropout_mask(uint8_t* mask, int32_t site, int32_t layer, int32_t rows, int32_t cols, float drop_p, uint64_t seed, void* stream) {
    TTTS_CHECK_ARG(site >= SITE_EMBD && site <= SITE_MLP_O && layer >= 0, "gpt_dropout_mask: bad site / layer");
    const DropCfg dc = site_drop(drop_p, seed, site, layer);
    if (site == SITE_ATTN_P) return attn_dropout_mask(mask, rows, cols, dc, (cudaStream_t)stream);
    return elem_dropout_mask(mask, rows, cols, dc, (cudaStream_t)stream);
}
int ttts_ce_fwd(const void* logits, int32_t ld, int32_t V, const int32_t* targets, int32_t rows, float* row_loss, float* row_lse, float* loss_out,
                void* stream) {
    return ce_fwd((const bf16*)logits, ld, V, targets, rows, row_loss, row_lse, loss_out, (cudaStream_t)stream);
}
int ttts_ce_bwd(const void* logits, int32_t ld, int32_t V, const int32_t* targets, int32_t rows, const float* row_lse, const float* gscale,
                float weight, void* dlogits, void* stream) {
    return ce_bwd((const bf16*)logits, ld, V, targets, rows, row_lse, gscale, weight, (bf16*)dlogits, (cudaStream_t)stream);
}
}
