"""CPU ORACLE (test infrastructure, never the product path) for the NEXT scope row: `enc_p_2`, the prior encoder of
SynthesizerTrn -- `TextEncoder(192, 192, 768, n_heads 2, n_layers 6, kernel 3)` (ttts/vqvae/vq2.py:101-164): a 3-layer relative-position
transformer over the up-sampled quantized latents, a 6-layer one over the embedded text, MRTE cross-attention between them with the style
vector added (vq2.py:17-50, MultiHeadAttention of utils/vc_utils.py:514-640), 3 more layers, and the projection to (m_p, logs_p).

The windowed relative-position attention of ttts/vqvae/attentions.py:177-363 (pad / reshape "skewing" tricks) is restated in closed form:
    scores[i, j] = q_i . k_j / sqrt(d) + [|j - i| <= w] q_i . Ek[j - i + w] / sqrt(d)
    out[i]       = sum_j p[i, j] v_j + sum_{|j - i| <= w} p[i, j] Ev[j - i + w]                                   (w = window_size = 4)
Eval mode (the p = 0.1 dropouts of the reference are off).  Pinned by tests/golden/make_golden.py::text_encoder_case against the REAL module
(tests/test_oracle_golden_text_encoder.py).  No kernels for this module yet."""
import math

import numpy as np
import torch
import torch.nn.functional as F

HID, FILT, HEADS, KS, WIN, OUT = 192, 768, 2, 3, 4, 192
MRTE_H, MRTE_HEADS = 512, 4


def _enc_shapes(s, pre, n_layers):
    dk = HID // HEADS
    for i in range(n_layers):
        for c in ("conv_q", "conv_k", "conv_v", "conv_o"):
            s[pre + "attn_layers.%d.%s.weight" % (i, c)] = (HID, HID, 1); s[pre + "attn_layers.%d.%s.bias" % (i, c)] = (HID,)
        s[pre + "attn_layers.%d.emb_rel_k" % i] = (1, 2 * WIN + 1, dk); s[pre + "attn_layers.%d.emb_rel_v" % i] = (1, 2 * WIN + 1, dk)
        for n in ("norm_layers_1", "norm_layers_2"):
            s[pre + "%s.%d.gamma" % (n, i)] = (HID,); s[pre + "%s.%d.beta" % (n, i)] = (HID,)
        s[pre + "ffn_layers.%d.conv_1.weight" % i] = (FILT, HID, KS); s[pre + "ffn_layers.%d.conv_1.bias" % i] = (FILT,)
        s[pre + "ffn_layers.%d.conv_2.weight" % i] = (HID, FILT, KS); s[pre + "ffn_layers.%d.conv_2.bias" % i] = (HID,)


def param_shapes():
    s = {}
    _enc_shapes(s, "encoder_ssl.", 3)
    _enc_shapes(s, "encoder_text.", 6)
    _enc_shapes(s, "encoder2.", 3)
    s["text_embedding.weight"] = (256, HID)
    for c in ("conv_q", "conv_k", "conv_v", "conv_o"):
        s["mrte.cross_attention.%s.weight" % c] = (MRTE_H, MRTE_H, 1); s["mrte.cross_attention.%s.bias" % c] = (MRTE_H,)
    s["mrte.c_pre.weight"] = (MRTE_H, HID, 1); s["mrte.c_pre.bias"] = (MRTE_H,)
    s["mrte.text_pre.weight"] = (MRTE_H, HID, 1); s["mrte.text_pre.bias"] = (MRTE_H,)
    s["mrte.c_post.weight"] = (HID, MRTE_H, 1); s["mrte.c_post.bias"] = (HID,)
    s["proj.weight"] = (2 * OUT, HID, 1); s["proj.bias"] = (2 * OUT,)
    return s


def init_params(seed=0):
    rs = np.random.RandomState(seed)
    out = {}
    shapes = param_shapes()
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("gamma"):
            v = rs.uniform(0.7, 1.3, size=shp)
        elif name.endswith("bias") or name.endswith("beta"):
            v = 0.05 * rs.standard_normal(shp)
        elif "emb_rel" in name:
            v = rs.standard_normal(shp) * shp[-1] ** -0.5
        elif name == "text_embedding.weight":
            v = rs.standard_normal(shp)
        else:
            v = rs.standard_normal(shp) / math.sqrt(int(np.prod(shp[1:])))
        out[name] = torch.tensor(v, dtype=torch.float32)
    return out


def layer_norm_c(x, gamma, beta):
    """modules.py:20-32: LayerNorm over the channel axis of [B, C, T]"""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, 1e-5).transpose(1, -1)


def _rel_table(emb, T):
    """[T, T, d]: emb[j - i + w] where |j - i| <= w, else 0 (attentions.py:310-347: the zero padding of _get_relative_embeddings)"""
    idx = torch.arange(T)[None, :] - torch.arange(T)[:, None]
    ok = idx.abs() <= WIN
    return emb[0][(idx + WIN).clamp(0, 2 * WIN)] * ok[:, :, None]


def self_attention(P, pre, x, attn_mask):
    """attentions.py:231-290 with window_size = 4; x [B, C, T], attn_mask [B, 1, T, T]"""
    B, C, T = x.shape
    dk = C // HEADS
    q = F.conv1d(x, P[pre + "conv_q.weight"], P[pre + "conv_q.bias"]).view(B, HEADS, dk, T).transpose(2, 3) / math.sqrt(dk)
    k = F.conv1d(x, P[pre + "conv_k.weight"], P[pre + "conv_k.bias"]).view(B, HEADS, dk, T).transpose(2, 3)
    v = F.conv1d(x, P[pre + "conv_v.weight"], P[pre + "conv_v.bias"]).view(B, HEADS, dk, T).transpose(2, 3)
    scores = q @ k.transpose(-2, -1) + torch.einsum("bhid,ijd->bhij", q, _rel_table(P[pre + "emb_rel_k"], T))
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    p = torch.softmax(scores, dim=-1)
    out = p @ v + torch.einsum("bhij,ijd->bhid", p, _rel_table(P[pre + "emb_rel_v"], T))
    out = out.transpose(2, 3).reshape(B, C, T)
    return F.conv1d(out, P[pre + "conv_o.weight"], P[pre + "conv_o.bias"])


def ffn(P, pre, x, x_mask):
    """attentions.py:406-414, kernel 3, "same" padding, relu"""
    h = torch.relu(F.conv1d(F.pad(x * x_mask, (1, 1)), P[pre + "conv_1.weight"], P[pre + "conv_1.bias"]))
    return F.conv1d(F.pad(h * x_mask, (1, 1)), P[pre + "conv_2.weight"], P[pre + "conv_2.bias"]) * x_mask


def encoder(P, pre, n_layers, x, x_mask):
    """attentions.py:66-88 (g = None)"""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    for i in range(n_layers):
        y = self_attention(P, pre + "attn_layers.%d." % i, x, attn_mask)
        x = layer_norm_c(x + y, P[pre + "norm_layers_1.%d.gamma" % i], P[pre + "norm_layers_1.%d.beta" % i])
        y = ffn(P, pre + "ffn_layers.%d." % i, x, x_mask)
        x = layer_norm_c(x + y, P[pre + "norm_layers_2.%d.gamma" % i], P[pre + "norm_layers_2.%d.beta" % i])
    return x * x_mask


def cross_attention(P, pre, x, c, attn_mask):
    """utils/vc_utils.py:568-620 without a window: queries from x [B, C, Tq], keys / values from c [B, C, Tk], attn_mask [B, 1, Tq, Tk]"""
    B, C, Tq = x.shape
    Tk = c.shape[2]
    dk = C // MRTE_HEADS
    q = F.conv1d(x, P[pre + "conv_q.weight"], P[pre + "conv_q.bias"]).view(B, MRTE_HEADS, dk, Tq).transpose(2, 3) / math.sqrt(dk)
    k = F.conv1d(c, P[pre + "conv_k.weight"], P[pre + "conv_k.bias"]).view(B, MRTE_HEADS, dk, Tk).transpose(2, 3)
    v = F.conv1d(c, P[pre + "conv_v.weight"], P[pre + "conv_v.bias"]).view(B, MRTE_HEADS, dk, Tk).transpose(2, 3)
    scores = (q @ k.transpose(-2, -1)).masked_fill(attn_mask == 0, -1e4)
    out = (torch.softmax(scores, dim=-1) @ v).transpose(2, 3).reshape(B, C, Tq)
    return F.conv1d(out, P[pre + "conv_o.weight"], P[pre + "conv_o.bias"])


def text_encoder(P, y, y_lengths, text, text_lengths, ge):
    """TextEncoder.forward (vq2.py:145-164): y [B, 192, T] quantized latents (up-sampled), text [B, Tt] int64, ge [B, 512, 1] -> (y, m, logs)"""
    y_mask = (torch.arange(y.shape[2])[None, :] < y_lengths[:, None]).unsqueeze(1).to(y.dtype)
    y = encoder(P, "encoder_ssl.", 3, y * y_mask, y_mask)
    text_mask = (torch.arange(text.shape[1])[None, :] < text_lengths[:, None]).unsqueeze(1).to(y.dtype)
    t = P["text_embedding.weight"][text].transpose(1, 2)
    t = encoder(P, "encoder_text.", 6, t * text_mask, text_mask)
    # MRTE (vq2.py:34-50)
    attn_mask = text_mask.unsqueeze(2) * y_mask.unsqueeze(-1)
    ssl = F.conv1d(y * y_mask, P["mrte.c_pre.weight"], P["mrte.c_pre.bias"])
    te = F.conv1d(t * text_mask, P["mrte.text_pre.weight"], P["mrte.text_pre.bias"])
    x = cross_attention(P, "mrte.cross_attention.", ssl * y_mask, te * text_mask, attn_mask) + ssl + ge
    y = F.conv1d(x * y_mask, P["mrte.c_post.weight"], P["mrte.c_post.bias"])
    y = encoder(P, "encoder2.", 3, y * y_mask, y_mask)
    stats = F.conv1d(y, P["proj.weight"], P["proj.bias"]) * y_mask
    m, logs = torch.split(stats, OUT, dim=1)
    return y, m, logs


def golden_inputs():
    g0 = torch.Generator().manual_seed(61)
    y = torch.randn(3, 192, 24, generator=g0)
    y_lengths = torch.tensor([24, 17, 6])
    text = torch.randint(0, 256, (3, 19), generator=g0)
    text_lengths = torch.tensor([19, 12, 3])
    ge = torch.randn(3, 512, 1, generator=g0)
    return y, y_lengths, text, text_lengths, ge
