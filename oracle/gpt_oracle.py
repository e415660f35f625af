"""CPU ORACLE (test infrastructure, never the product path) for the UnifiedVoice GPT train step.

A plain-torch fp32 restatement, with no dependency on HuggingFace `transformers`, of exactly what the
reference executes on its hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.

Parity status: the reference ships NO tests or golden vectors ("parity unpinned by the reference's own
tests").  This restatement is pinned instead against the REAL reference modules
executed in the build container: tests/golden/make_golden.py imports /root/reference/ttts/gpt/model.py
(with the import shims of SURVEY.md Appendix D), runs it in fp32 eval mode on seeded inputs and commits
inputs + outputs under tests/golden/; tests/test_oracle_golden.py checks this file against those vectors.

Reference citations (paths under /root/reference unless prefixed HF: = transformers 5.5.0):
  pre-processing ............. ttts/gpt/model.py:471-489  (clip, set_mel_padding 402-414, start/stop padding 397-400)
  embeddings + learned pos ... ttts/gpt/model.py:230-242, 488, 494-495, 418
  GPT-2 block ................ HF: models/gpt2/modeling_gpt2.py:246-310 (block), 144-226 (attention), 229-243 (MLP),
                               pytorch_utils.py:97-123 (Conv1D: y = x @ W + b, W stored [in,out]),
                               activations.py:59-66 (gelu_new)
  wpe == 0 ................... ttts/gpt/model.py:12-13, 260-261
  ln_f, final_norm, heads .... HF: modeling_gpt2.py:628 ; ttts/gpt/model.py:426-438
  losses ..................... ttts/gpt/model.py:508-510 (mean CE over ALL positions, no ignore_index)
  step tail .................. ttts/gpt/train.py:109-120 (0.01*loss_text + loss_mel, clip 1.0, AdamW(0.9,0.96, wd .01))
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-5


def default_config(**over):
    cfg = dict(layers=6, model_dim=512, heads=8, max_text_tokens=800, max_mel_tokens=1600, number_text_tokens=256,
               start_text_token=255, number_mel_codes=1026, start_mel_token=1024, stop_mel_token=1025,
               mel_length_compression=1024, types=1)
    cfg.update(over)
    return cfg


def param_shapes(cfg):
    """state_dict keys and shapes of the reference UnifiedVoice, in state_dict order."""
    d, L = cfg["model_dim"], cfg["layers"]
    vt = cfg["number_text_tokens"] * cfg.get("types", 1) + 1
    vm = cfg["number_mel_codes"]
    out = [("text_embedding.weight", (vt, d)), ("mel_embedding.weight", (vm, d))]
    for i in range(L):
        p = "gpt.h.%d." % i
        out += [(p + "ln_1.weight", (d,)), (p + "ln_1.bias", (d,)),
                (p + "attn.c_attn.weight", (d, 3 * d)), (p + "attn.c_attn.bias", (3 * d,)),
                (p + "attn.c_proj.weight", (d, d)), (p + "attn.c_proj.bias", (d,)),
                (p + "ln_2.weight", (d,)), (p + "ln_2.bias", (d,)),
                (p + "mlp.c_fc.weight", (d, 4 * d)), (p + "mlp.c_fc.bias", (4 * d,)),
                (p + "mlp.c_proj.weight", (4 * d, d)), (p + "mlp.c_proj.bias", (d,))]
    out += [("gpt.ln_f.weight", (d,)), ("gpt.ln_f.bias", (d,)),
            ("mel_pos_embedding.emb.weight", (cfg["max_mel_tokens"] + 2, d)),
            ("text_pos_embedding.emb.weight", (cfg["max_text_tokens"] + 2, d)),
            ("final_norm.weight", (d,)), ("final_norm.bias", (d,)),
            ("text_head.weight", (vt, d)), ("text_head.bias", (vt,)),
            ("mel_head.weight", (vm, d)), ("mel_head.bias", (vm,))]
    return out


def init_params(cfg, seed=0, dtype=torch.float32):
    """Deterministic (numpy-seeded, torch-version independent) parameters with the reference's init statistics."""
    import numpy as np
    rs = np.random.RandomState(seed)
    L = cfg["layers"]
    params = {}
    for name, shape in param_shapes(cfg):
        if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name in ("gpt.ln_f.weight", "final_norm.weight"):
            # perturbed away from 1/0 so LayerNorm affine gradients are exercised
            v = 1.0 + 0.05 * rs.standard_normal(shape)
        elif name.endswith(".bias"):
            v = 0.02 * rs.standard_normal(shape)
        elif name.endswith("c_proj.weight"):
            v = (0.02 / math.sqrt(2 * L)) * rs.standard_normal(shape)
        elif name.startswith("text_head") or name.startswith("mel_head"):
            bound = 1.0 / math.sqrt(shape[1])
            v = rs.uniform(-bound, bound, size=shape)
        else:
            v = 0.02 * rs.standard_normal(shape)
        params[name] = torch.tensor(v, dtype=dtype)
    return params


def gelu_new(x):
    # HF: activations.py:59-66
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


def _mm(a, b, emulate_bf16):
    if emulate_bf16:
        return (a.to(torch.bfloat16).float() @ b.to(torch.bfloat16).float()).to(torch.bfloat16).float()
    return a @ b


def preprocess(cfg, text_inputs, text_lengths, mel_codes, wav_lengths, clip_inputs=True):
    """ttts/gpt/model.py:471-489.  NOTE: like the reference, mutates `mel_codes` in place (set_mel_padding)."""
    stop_text, start_text = 0, cfg["start_text_token"]
    start_mel, stop_mel = cfg["start_mel_token"], cfg["stop_mel_token"]
    mlc = cfg.get("mel_length_compression", 1024)
    if clip_inputs:
        max_text_len = int(text_lengths.max())
        text_inputs = text_inputs[:, :max_text_len]
        max_mel_len = int(wav_lengths.max()) // mlc
        mel_codes = mel_codes[:, :max_mel_len]
    mel_lengths = torch.div(wav_lengths, mlc, rounding_mode="trunc")
    for b in range(len(mel_lengths)):
        actual_end = int(mel_lengths[b]) + 1
        if actual_end < mel_codes.shape[-1]:
            mel_codes[b, actual_end:] = stop_mel
    text_inputs = F.pad(text_inputs, (0, 1), value=stop_text)
    mel_codes = F.pad(mel_codes, (0, 1), value=stop_mel)
    text_in = F.pad(text_inputs, (1, 0), value=start_text)
    text_tgt = F.pad(text_inputs, (0, 1), value=stop_text)
    mel_in = F.pad(mel_codes, (1, 0), value=start_mel)
    mel_tgt = F.pad(mel_codes, (0, 1), value=stop_mel)
    return text_in, text_tgt, mel_in, mel_tgt


def gpt_hidden(params, cfg, text_in, mel_in, emulate_bf16=False, collect=None, masks=None, drop_scale=1.0):
    """Embeddings -> L GPT-2 blocks -> ln_f -> final_norm.  Returns enc (B,T,d) fp32.
    masks (training mode, HF GPT2Config defaults embd_pdrop = attn_pdrop = resid_pdrop = 0.1): dict of 0/1 keep masks "embd" [B,T,d],
    "attn_p%d" [B,H,T,T], "attn_o%d" / "mlp_o%d" [B,T,d]; kept elements are scaled by drop_scale = 1/(1-p), as torch's dropout does
    (HF:modeling_gpt2.py GPT2Model.drop, GPT2Attention.attn_dropout / resid_dropout, GPT2MLP.dropout).  The mask VALUES are an input: the
    CUDA path draws them from its own counter-based generator and exports them (ttts_gpt_dropout_mask)."""
    dm = (lambda t, key: t * masks[key].to(t.dtype) * drop_scale) if masks is not None else (lambda t, key: t)
    d, H = cfg["model_dim"], cfg["heads"]
    hd = d // H
    Tt, Tm = text_in.shape[1], mel_in.shape[1]
    text_emb = params["text_embedding.weight"][text_in] + params["text_pos_embedding.emb.weight"][:Tt]
    mel_emb = params["mel_embedding.weight"][mel_in] + params["mel_pos_embedding.emb.weight"][:Tm]
    x = torch.cat([text_emb, mel_emb], dim=1)          # fp32 residual stream (Appendix A)
    x = dm(x, "embd")
    B, T, _ = x.shape
    causal = torch.ones(T, T, dtype=torch.bool, device=x.device).tril()
    r = (lambda t: t.to(torch.bfloat16).float()) if emulate_bf16 else (lambda t: t)
    if collect is not None:
        collect["x0"] = x
    for i in range(cfg["layers"]):
        p = "gpt.h.%d." % i
        h = layer_norm(x, params[p + "ln_1.weight"], params[p + "ln_1.bias"])
        qkv = _mm(h, params[p + "attn.c_attn.weight"], emulate_bf16)
        qkv = r(qkv + params[p + "attn.c_attn.bias"]) if emulate_bf16 else qkv + params[p + "attn.c_attn.bias"]
        q, k, v = qkv.split(d, dim=2)
        q = q.view(B, T, H, hd).transpose(1, 2)
        k = k.view(B, T, H, hd).transpose(1, 2)
        v = v.view(B, T, H, hd).transpose(1, 2)
        if collect is not None:
            collect["k%d" % i], collect["v%d" % i] = k, v
        att = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
        att = att.masked_fill(~causal, float("-inf"))
        att = torch.softmax(att, dim=-1)
        att = dm(att, "attn_p%d" % i)
        a = r(r(att) @ v) if emulate_bf16 else att @ v
        a = a.transpose(1, 2).reshape(B, T, d)
        o = _mm(a, params[p + "attn.c_proj.weight"], emulate_bf16)
        o = r(o + params[p + "attn.c_proj.bias"]) if emulate_bf16 else o + params[p + "attn.c_proj.bias"]
        x = x + dm(o, "attn_o%d" % i)
        h = layer_norm(x, params[p + "ln_2.weight"], params[p + "ln_2.bias"])
        f = _mm(h, params[p + "mlp.c_fc.weight"], emulate_bf16)
        f = r(f + params[p + "mlp.c_fc.bias"]) if emulate_bf16 else f + params[p + "mlp.c_fc.bias"]
        g = r(gelu_new(f)) if emulate_bf16 else gelu_new(f)
        o = _mm(g, params[p + "mlp.c_proj.weight"], emulate_bf16)
        o = r(o + params[p + "mlp.c_proj.bias"]) if emulate_bf16 else o + params[p + "mlp.c_proj.bias"]
        x = x + dm(o, "mlp_o%d" % i)
        if collect is not None:
            collect["x%d" % (i + 1)] = x
    x = layer_norm(x, params["gpt.ln_f.weight"], params["gpt.ln_f.bias"])
    x = layer_norm(x, params["final_norm.weight"], params["final_norm.bias"])
    return x


def forward(params, cfg, text_inputs, text_lengths, mel_codes, wav_lengths, clip_inputs=True, return_latent=False,
            emulate_bf16=False, collect=None, masks=None, drop_scale=1.0):
    """UnifiedVoice.forward (ttts/gpt/model.py:453-510), text_first=True; eval mode (dropout off) unless `masks` is given (gpt_hidden)."""
    text_in, text_tgt, mel_in, mel_tgt = preprocess(cfg, text_inputs, text_lengths, mel_codes, wav_lengths, clip_inputs)
    enc = gpt_hidden(params, cfg, text_in, mel_in, emulate_bf16, collect, masks, drop_scale)
    Tt, Tm = text_in.shape[1], mel_in.shape[1]
    if return_latent:
        return enc[:, -Tm:][:, :-2]
    r = (lambda t: t.to(torch.bfloat16).float()) if emulate_bf16 else (lambda t: t)
    text_logits = _mm(enc[:, :Tt], params["text_head.weight"].t(), emulate_bf16)
    text_logits = r(text_logits + params["text_head.bias"])
    mel_logits = _mm(enc[:, -Tm:], params["mel_head.weight"].t(), emulate_bf16)
    mel_logits = r(mel_logits + params["mel_head.bias"])
    text_logits = text_logits.permute(0, 2, 1)
    mel_logits = mel_logits.permute(0, 2, 1)
    loss_text = F.cross_entropy(text_logits, text_tgt.long())
    loss_mel = F.cross_entropy(mel_logits, mel_tgt.long())
    return loss_text.mean(), loss_mel.mean(), mel_logits


# ---------------------------------------------------------------------------------------------------------------------
# KV-cache decode: the cached branch of GPT2InferenceModel.forward (ttts/gpt/model.py:106-171) as a plain incremental
# restatement.  The cache has the layout the CUDA path uses, [layers, 2 (k|v), B, heads, T_max, head_dim].
# ---------------------------------------------------------------------------------------------------------------------
def kv_prefill(params, cfg, text_in, mel_in, T_max, emulate_bf16=False):
    """Full forward over the prompt [text_in ; mel_in] (both already start/stop padded, mel_in = [start_mel, codes...]).
    Returns (cache, slot, logits of the last position [B, V_mel]); slot = number of cached positions."""
    col = {}
    enc = gpt_hidden(params, cfg, text_in, mel_in, emulate_bf16, col)
    B, T = enc.shape[0], enc.shape[1]
    H = cfg["heads"]
    hd = cfg["model_dim"] // H
    assert T <= T_max
    cache = torch.zeros(cfg["layers"], 2, B, H, T_max, hd, dtype=enc.dtype)
    for i in range(cfg["layers"]):
        cache[i, 0, :, :, :T] = col["k%d" % i]
        cache[i, 1, :, :, :T] = col["v%d" % i]
    r = (lambda t: t.to(torch.bfloat16).float()) if emulate_bf16 else (lambda t: t)
    logits = r(_mm(enc[:, -1], params["mel_head.weight"].t(), emulate_bf16) + params["mel_head.bias"])
    return cache, T, logits


def kv_decode_step(params, cfg, cache, slot, tokens, text_positions, pos_shift=0, emulate_bf16=False):
    """One cached step: `tokens` [B] int64 are the codes being fed, stored at cache slot `slot`.  Position embedding index of the token:
    its index inside the mel segment (`slot - text_positions`, what the uncached path uses, model.py:134-142) + pos_shift; the reference's
    cached branch uses `attention_mask.shape[1] - mel_len` = that index + 1 (model.py:144-147), i.e. pos_shift = 1.
    Returns logits [B, V_mel]; the cache is updated in place (caller advances slot)."""
    d, H = cfg["model_dim"], cfg["heads"]
    hd = d // H
    B = tokens.shape[0]
    r = (lambda t: t.to(torch.bfloat16).float()) if emulate_bf16 else (lambda t: t)
    pos = slot - text_positions + pos_shift
    x = params["mel_embedding.weight"][tokens] + params["mel_pos_embedding.emb.weight"][pos]
    for i in range(cfg["layers"]):
        p = "gpt.h.%d." % i
        h = layer_norm(x, params[p + "ln_1.weight"], params[p + "ln_1.bias"])
        qkv = r(_mm(h, params[p + "attn.c_attn.weight"], emulate_bf16) + params[p + "attn.c_attn.bias"])
        q, k, v = qkv.split(d, dim=1)
        cache[i, 0, :, :, slot] = k.view(B, H, hd)
        cache[i, 1, :, :, slot] = v.view(B, H, hd)
        K, V = cache[i, 0, :, :, :slot + 1], cache[i, 1, :, :, :slot + 1]                # [B, H, n, hd]
        att = torch.einsum("bhd,bhnd->bhn", q.view(B, H, hd), K) * (hd ** -0.5)
        att = torch.softmax(att, dim=-1)
        a = r(torch.einsum("bhn,bhnd->bhd", r(att), V)).reshape(B, d)
        x = x + r(_mm(a, params[p + "attn.c_proj.weight"], emulate_bf16) + params[p + "attn.c_proj.bias"])
        h = layer_norm(x, params[p + "ln_2.weight"], params[p + "ln_2.bias"])
        f = r(_mm(h, params[p + "mlp.c_fc.weight"], emulate_bf16) + params[p + "mlp.c_fc.bias"])
        g = r(gelu_new(f))
        x = x + r(_mm(g, params[p + "mlp.c_proj.weight"], emulate_bf16) + params[p + "mlp.c_proj.bias"])
    x = layer_norm(x, params["gpt.ln_f.weight"], params["gpt.ln_f.bias"])
    x = layer_norm(x, params["final_norm.weight"], params["final_norm.bias"])
    return r(_mm(x, params["mel_head.weight"].t(), emulate_bf16) + params["mel_head.bias"])


def loss_and_grads(params, cfg, text_inputs, text_lengths, mel_codes, wav_lengths, text_weight=0.01, mel_weight=1.0,
                   emulate_bf16=False, masks=None, drop_scale=1.0):
    """fwd + bwd of loss = text_weight*loss_text + mel_weight*loss_mel (ttts/gpt/train.py:109-112)."""
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    lt, lm, logits = forward(ps, cfg, text_inputs, text_lengths, mel_codes.clone(), wav_lengths, emulate_bf16=emulate_bf16, masks=masks,
                             drop_scale=drop_scale)
    loss = lt * text_weight + lm * mel_weight
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in ps.items()}
    return lt.detach(), lm.detach(), logits.detach(), grads


def clip_and_adamw(params, grads, state, lr, step, max_norm=1.0, betas=(0.9, 0.96), eps=1e-8, weight_decay=0.01):
    """ttts/gpt/train.py:114-118: global-norm clip (torch clip_grad_norm_ semantics) then torch.optim.AdamW.
    `state` maps name -> (exp_avg, exp_avg_sq); `step` is the 1-based optimizer step.  Updates in place."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k] * coef
        if k not in state:
            state[k] = (torch.zeros_like(p), torch.zeros_like(p))
        m, v = state[k]
        p.mul_(1.0 - lr * weight_decay)
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        bc1 = 1.0 - b1 ** step
        bc2 = 1.0 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)
    return total


def warmup_lr(step, base_lr=1e-4):
    """ttts/gpt/train.py:36-40 LambdaLR: lr = base * step/500 for step < 500 (scheduler step index, 0-based)."""
    return base_lr * (float(step) / 500.0 if step < 500 else 1.0)


def synthetic_batch(B, TL, CL, seed=1234, device="cpu"):
    """BASELINE.md section 3 synthetic inputs."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(1, 255, (B, TL), generator=g, dtype=torch.int64)
    codes = torch.randint(0, 1024, (B, CL), generator=g, dtype=torch.int64)
    text_lengths = torch.full((B,), TL, dtype=torch.int64)
    wav_lengths = torch.full((B,), CL * 1024, dtype=torch.int64)
    return text.to(device), text_lengths.to(device), codes.to(device), wav_lengths.to(device)


def flops_per_step(cfg, B, TL, CL):
    """SURVEY.md section 8d algorithmic FLOP model (causal attention counted half, no recompute credit)."""
    d, L = cfg["model_dim"], cfg["layers"]
    T = TL + CL + 4
    vt = cfg["number_text_tokens"] * cfg.get("types", 1) + 1
    vm = cfg["number_mel_codes"]
    fwd = L * (24 * d * d * T + 2 * T * T * d) + 2 * d * vm * (CL + 2) + 2 * d * vt * (TL + 2)
    return 3 * B * fwd


# ---------------------------------------------------------------------------------------------------------------------
# Attention-probability dropout mask (HF: modeling_gpt2.py:216 `attn_dropout`).  torch's Philox stream cannot be matched by a
# fused kernel, so the CUDA path defines its own counter-based keep function (ttts_b200/csrc/common.cuh attn_drop_*); this is
# its numpy restatement.  With the mask in hand, plain torch reproduces the dropped attention exactly, and the statistical
# quality of the hash (keep rate, correlations, spectrum) is checked on the CPU in tests/test_oracle_golden.py.
# ---------------------------------------------------------------------------------------------------------------------
def attn_dropout_thresh16(p):
    """p rounded to a multiple of 2^-15, as a 16-bit threshold (always even): the attention hash decides on 15-bit fields"""
    return 2 * int(p * 32768.0 + 0.5)


def attn_dropout_keep_mask(seed, rows, T, p):
    """keep[r, kj] for mask rows `rows` (= (b*H + h)*T + query index, any integer array) and keys 0..T-1; bool [len(rows), T]."""
    import numpy as np
    U = np.uint64
    lo32 = U(0xFFFFFFFF)
    rows = np.asarray(rows, dtype=U)
    with np.errstate(over="ignore"):
        z = U(seed) + rows * U(0x9E3779B97F4A7C15)
        z ^= z >> U(33); z *= U(0xFF51AFD7ED558CCD)
        z ^= z >> U(33); z *= U(0xC4CEB9FE1A85EC53)
        z ^= z >> U(33)
        k0, k1 = (z & lo32)[:, None], (z >> U(32))[:, None]
        g = np.arange((T + 3) // 4, dtype=U)[None, :]
        a = (g * U(0x9E3779B1) + k0) & lo32
        m1 = a * U(0x85EBCA6B)
        x = (m1 & lo32) ^ (m1 >> U(32)) ^ k1
        m2 = x * U(0xC2B2AE35)
        y = (m2 & lo32) ^ (m2 >> U(32))
        m3 = y * U(0x27D4EB2F)
        w0 = (m3 >> U(32)) ^ (m2 & lo32)
        w1 = (m3 & lo32) ^ (m2 >> U(32))
        t15 = U(attn_dropout_thresh16(p) >> 1)
        m15 = U(0x7FFF)
        # four 15-bit fields per group: key 4g+0 <- w0 bits 0-14, 4g+1 <- w0 bits 16-30, 4g+2 <- w1 bits 0-14, 4g+3 <- w1 bits 16-30
        f = np.stack([w0 & m15, (w0 >> U(16)) & m15, w1 & m15, (w1 >> U(16)) & m15], axis=-1)
    return (f >= t15).reshape(len(rows), -1)[:, :T]
