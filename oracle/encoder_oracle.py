"""CPU ORACLE (test infrastructure, never the product path) for the VQ-VAE encoder stack: MelStyleEncoder (ref_enc),
PosteriorAudioEncoder (enc_p: strided weight-normed convs, 15 ResBlock1, anti-aliased SnakeBeta, WN with global conditioning),
the stride-2 `proj` and the quantizer lookup -- i.e. the encode half of SynthesizerTrn.  Plain torch functional ops on a
name -> tensor dict that uses the REFERENCE's state_dict names.

Parity status: the reference has no tests ("parity unpinned"); pinned instead against the real reference modules
run in the build container (tests/golden/make_golden.py -> tests/golden/encoder.npz; tests/test_oracle_golden_encoder.py).

Reference citations (under /root/reference/ttts/vqvae):
  MelStyleEncoder ........ modules.py:686-764 (LinearNorm 529-547, Mish 550-555, Conv1dGLU 558-567, MultiHeadAttention 600-654,
                           ScaledDotProductAttention 657-676)
  PosteriorAudioEncoder .. vq2.py:667-745 ; ResBlock1 modules.py:224-318 ; WN modules.py:136-221 ;
                           fused_add_tanh_sigmoid_multiply ttts/utils/commons.py:102-109
  Activation1d/SnakeBeta . alias_free_torch/act.py:8-28, resample.py:11-49, filter.py:29-95 ; activations.py:62-119
  encode pipeline ........ vq2.py:843-852 / 874-882 ; spectrogram_torch ttts/utils/data_utils.py:52-87
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

CH = [16, 32, 64, 96, 128, 192]
RATES = [10, 8, 2, 2, 2]
KSZ = [16, 16, 8, 2, 2]
HID, GIN, SPEC = 192, 512, 1025


def param_shapes():
    """name -> shape for the encode half of SynthesizerTrn, in the reference's state_dict naming."""
    s = {}
    # ref_enc = MelStyleEncoder(1025, style_vector_dim=512)
    s["ref_enc.spectral.0.fc.weight"] = (128, SPEC); s["ref_enc.spectral.0.fc.bias"] = (128,)
    s["ref_enc.spectral.3.fc.weight"] = (128, 128); s["ref_enc.spectral.3.fc.bias"] = (128,)
    for i in range(2):
        s["ref_enc.temporal.%d.conv1.conv.weight" % i] = (256, 128, 5); s["ref_enc.temporal.%d.conv1.conv.bias" % i] = (256,)
    for n in ("w_qs", "w_ks", "w_vs", "fc"):
        s["ref_enc.slf_attn.%s.weight" % n] = (128, 128); s["ref_enc.slf_attn.%s.bias" % n] = (128,)
    s["ref_enc.fc.fc.weight"] = (GIN, 128); s["ref_enc.fc.fc.bias"] = (GIN,)
    # enc_p = PosteriorAudioEncoder(1025, 192, 192, 5, 1, 16, gin_channels=512)
    s["enc_p.pre.weight"] = (HID, SPEC, 1); s["enc_p.pre.bias"] = (HID,)
    s["enc_p.down_pre.weight"] = (16, 1, 7); s["enc_p.down_pre.bias"] = (16,)
    for i in range(5):
        s["enc_p.downs.%d.weight_g" % i] = (CH[i + 1], 1, 1); s["enc_p.downs.%d.weight_v" % i] = (CH[i + 1], CH[i], KSZ[i])
        s["enc_p.downs.%d.bias" % i] = (CH[i + 1],)
    for i in range(5):
        for j, k in enumerate((3, 7, 11)):
            c = CH[i + 1]
            for cs in ("convs1", "convs2"):
                for t in range(3):
                    p = "enc_p.resblocks.%d.%s.%d." % (i * 3 + j, cs, t)
                    s[p + "parametrizations.weight.original0"] = (c, 1, 1)
                    s[p + "parametrizations.weight.original1"] = (c, c, k)
                    s[p + "bias"] = (c,)
    s["enc_p.activation_post.act.alpha"] = (192,); s["enc_p.activation_post.act.beta"] = (192,)
    s["enc_p.conv_post.weight"] = (HID, 192, 7); s["enc_p.conv_post.bias"] = (HID,)
    s["enc_p.enc.cond_layer.weight_g"] = (2 * HID * 16, 1, 1); s["enc_p.enc.cond_layer.weight_v"] = (2 * HID * 16, GIN, 1)
    s["enc_p.enc.cond_layer.bias"] = (2 * HID * 16,)
    for i in range(16):
        s["enc_p.enc.in_layers.%d.weight_g" % i] = (2 * HID, 1, 1); s["enc_p.enc.in_layers.%d.weight_v" % i] = (2 * HID, HID, 5)
        s["enc_p.enc.in_layers.%d.bias" % i] = (2 * HID,)
        co = 2 * HID if i < 15 else HID
        s["enc_p.enc.res_skip_layers.%d.weight_g" % i] = (co, 1, 1); s["enc_p.enc.res_skip_layers.%d.weight_v" % i] = (co, HID, 1)
        s["enc_p.enc.res_skip_layers.%d.bias" % i] = (co,)
    s["enc_p.proj.weight"] = (2 * HID, 2 * HID, 1); s["enc_p.proj.bias"] = (2 * HID,)
    s["proj.weight"] = (HID, HID, 2); s["proj.bias"] = (HID,)
    return s


def init_params(seed=0):
    """Deterministic (numpy-seeded) parameters with magnitudes that keep activations O(1) through the stack."""
    rs = np.random.RandomState(seed)
    out = {}
    shapes = param_shapes()
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("weight_g") or name.endswith("original0"):
            v = rs.uniform(0.6, 1.4, size=shp)
        elif name.endswith(".alpha") or name.endswith(".beta"):
            v = 0.3 * rs.standard_normal(shp)
        elif name.endswith("bias"):
            v = 0.05 * rs.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            v = rs.standard_normal(shp) / math.sqrt(fan_in)
        out[name] = torch.tensor(v, dtype=torch.float32)
    # weight-norm g chosen relative to ||v|| so effective weights have ~unit gain
    for name in list(out):
        if name.endswith("weight_g") or name.endswith("original0"):
            vname = name[:-len("weight_g")] + "weight_v" if name.endswith("weight_g") else name[:-1] + "1"
            vn = out[vname].flatten(1).norm(dim=1).view(out[name].shape)
            out[name] = out[name] * vn * 0.9
    return out


def kaiser_sinc_filter12():
    cutoff, half_width, ks = 0.25, 0.3, 12
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    window = torch.kaiser_window(ks, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (f / f.sum()).view(1, 1, ks)


def _wn(P, prefix, old=True):
    if old:
        g, v = P[prefix + "weight_g"], P[prefix + "weight_v"]
    else:
        g, v = P[prefix + "parametrizations.weight.original0"], P[prefix + "parametrizations.weight.original1"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def mel_style_encoder(P, x, mask):
    """x [B,1025,T] (pre-masked), mask [B,1,T] float."""
    pre = "ref_enc."
    B, _, T = x.shape
    pad = (mask.int() == 0).squeeze(1)                                  # True at padding
    h = x.transpose(1, 2)
    mish = lambda t: t * torch.tanh(F.softplus(t))
    h = mish(F.linear(h, P[pre + "spectral.0.fc.weight"], P[pre + "spectral.0.fc.bias"]))
    h = mish(F.linear(h, P[pre + "spectral.3.fc.weight"], P[pre + "spectral.3.fc.bias"]))
    h = h.transpose(1, 2)
    for i in range(2):
        c = F.conv1d(h, P[pre + "temporal.%d.conv1.conv.weight" % i], P[pre + "temporal.%d.conv1.conv.bias" % i], padding=2)
        a, g = torch.split(c, 128, dim=1)
        h = h + a * torch.sigmoid(g)
    h = h.transpose(1, 2).masked_fill(pad.unsqueeze(-1), 0)
    H, dk = 2, 64
    q = F.linear(h, P[pre + "slf_attn.w_qs.weight"], P[pre + "slf_attn.w_qs.bias"]).view(B, T, H, dk).permute(2, 0, 1, 3).reshape(-1, T, dk)
    k = F.linear(h, P[pre + "slf_attn.w_ks.weight"], P[pre + "slf_attn.w_ks.bias"]).view(B, T, H, dk).permute(2, 0, 1, 3).reshape(-1, T, dk)
    v = F.linear(h, P[pre + "slf_attn.w_vs.weight"], P[pre + "slf_attn.w_vs.bias"]).view(B, T, H, dk).permute(2, 0, 1, 3).reshape(-1, T, dk)
    att = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(128.0)
    att = att.masked_fill(pad.unsqueeze(1).expand(-1, T, -1).repeat(H, 1, 1), -np.inf)
    o = torch.bmm(torch.softmax(att, dim=2), v).view(H, B, T, dk).permute(1, 2, 0, 3).reshape(B, T, -1)
    h = F.linear(o, P[pre + "slf_attn.fc.weight"], P[pre + "slf_attn.fc.bias"]) + h
    h = F.linear(h, P[pre + "fc.fc.weight"], P[pre + "fc.fc.bias"])
    lens = (~pad).sum(dim=1).unsqueeze(1)
    return (h.masked_fill(pad.unsqueeze(-1), 0).sum(dim=1) / lens).unsqueeze(-1)


def resblock1(P, prefix, x, k):
    for t, d in enumerate((1, 3, 5)):
        xt = F.leaky_relu(x, 0.1)
        xt = F.conv1d(xt, _wn(P, prefix + "convs1.%d." % t, old=False), P[prefix + "convs1.%d.bias" % t], dilation=d, padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, 0.1)
        xt = F.conv1d(xt, _wn(P, prefix + "convs2.%d." % t, old=False), P[prefix + "convs2.%d.bias" % t], padding=(k - 1) // 2)
        x = xt + x
    return x


def activation1d_snakebeta(P, x):
    C = x.shape[1]
    f = kaiser_sinc_filter12()
    xp = F.pad(x, (5, 5), mode="replicate")
    up = 2 * F.conv_transpose1d(xp, f.expand(C, -1, -1), stride=2, groups=C)[..., 15:-15]
    alpha = torch.exp(P["enc_p.activation_post.act.alpha"]).view(1, -1, 1)
    beta = torch.exp(P["enc_p.activation_post.act.beta"]).view(1, -1, 1)
    up = up + (1.0 / (beta + 1e-9)) * torch.sin(up * alpha) ** 2
    dp = F.pad(up, (5, 6), mode="replicate")
    return F.conv1d(dp, f.expand(C, -1, -1), stride=2, groups=C)


def wn(P, x, x_mask, g):
    pre = "enc_p.enc."
    out = torch.zeros_like(x)
    gc = F.conv1d(g, _wn(P, pre + "cond_layer."), P[pre + "cond_layer.bias"])
    for i in range(16):
        x_in = F.conv1d(x, _wn(P, pre + "in_layers.%d." % i), P[pre + "in_layers.%d.bias" % i], padding=2)
        g_l = gc[:, i * 2 * HID:(i + 1) * 2 * HID, :]
        a = x_in + g_l
        acts = torch.tanh(a[:, :HID]) * torch.sigmoid(a[:, HID:])
        rs = F.conv1d(acts, _wn(P, pre + "res_skip_layers.%d." % i), P[pre + "res_skip_layers.%d.bias" % i])
        if i < 15:
            x = (x + rs[:, :HID]) * x_mask
            out = out + rs[:, HID:]
        else:
            out = out + rs
    return out * x_mask


def posterior_audio_encoder(P, spec, wav, x_mask, g, eps=None):
    a = F.conv1d(wav, P["enc_p.down_pre.weight"], P["enc_p.down_pre.bias"], padding=3)
    for i in range(5):
        a = F.conv1d(a, _wn(P, "enc_p.downs.%d." % i), P["enc_p.downs.%d.bias" % i], stride=RATES[i], padding=(KSZ[i] - 1) // 2)
        xs = None
        for j, k in enumerate((3, 7, 11)):
            r = resblock1(P, "enc_p.resblocks.%d." % (i * 3 + j), a, k)
            xs = r if xs is None else xs + r
        a = xs / 3
    a = activation1d_snakebeta(P, a)
    a = F.conv1d(a, P["enc_p.conv_post.weight"], P["enc_p.conv_post.bias"], padding=3)
    x = F.conv1d(spec, P["enc_p.pre.weight"], P["enc_p.pre.bias"]) * x_mask
    x = wn(P, x, x_mask, g)
    a = a * x_mask
    stats = F.conv1d(torch.cat([x, a], dim=1), P["enc_p.proj.weight"], P["enc_p.proj.bias"]) * x_mask
    m, logs = torch.split(stats, HID, dim=1)
    e = eps if eps is not None else torch.zeros_like(m)
    return (m + e * torch.exp(logs)) * x_mask, m, logs


def encode(P, spec, wav, lengths=None, eps=None, codebook=None):
    """spec [B,1025,T], wav [B,L] -> dict(ge, z, m, logs, x, codes)."""
    from . import vq_mel_oracle as V
    B, _, T = spec.shape
    if lengths is None:
        mask = torch.ones(B, 1, T)
    else:
        mask = (torch.arange(T)[None, :] < lengths[:, None]).float().unsqueeze(1)
    ge = mel_style_encoder(P, spec * mask, mask)
    z, m, logs = posterior_audio_encoder(P, spec, wav.unsqueeze(1), mask, ge, eps)
    x = F.conv1d(z, P["proj.weight"], P["proj.bias"], stride=2)
    out = dict(ge=ge, z=z, m=m, logs=logs, x=x)
    if codebook is not None:
        xn = np.ascontiguousarray(x.numpy().transpose(0, 2, 1)).reshape(-1, x.shape[1])
        out["codes"] = V.vq_quantize(xn, codebook).reshape(1, B, -1)
    return out
