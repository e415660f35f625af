"""CPU ORACLE (test infrastructure, never the product path) for SURVEY.md 8(f) #3 -- the diffusion mel-refiner train step, BASELINE config 5:
`AA_diffusion` (ttts/diffusion/aa_model.py:182-287) under `SpacedDiffusion.training_losses` (ttts/utils/diffusion.py:930-1014, built at
ttts/diffusion/train.py:90-92: 1000 linear-beta steps, epsilon prediction, learned-range variance, MSE + variational-bound term).

Plain functional torch, fp32, on the reference's state_dict names.  Restated pieces:
  * GroupNorm32 / `normalization` (ttts/utils/utils.py:119-137), `AttentionBlock` + `QKVAttentionLegacy` (utils.py:140-215: heads are split
    BEFORE q/k/v, both q and k scaled by ch^-1/4, non-causal softmax) with the T5-style `RelativePositionBias` (utils/xtransformers.py:146-188:
    32 buckets, max_distance 64, bias * sqrt(ch) added to the scores);
  * `ResBlock` with scale-shift norm and the 1x1 "efficient" input convolution (aa_model.py:69-135), `DiffusionLayer` (:138-151),
    `RefEncoder` (:153-177; vc_utils.MultiHeadAttention cross attention of 32 learned latents over the reference mel), `timestep_embedding`;
  * the model's three random decisions are INPUTS here: the per-sample unconditioned mask (aa_model.py:247-252), the dropped layers
    (:269-271) and, in the loss, t and the noise (train.py:170, diffusion.py:945-947);
  * `q_sample`, `q_posterior_mean_variance`, `p_mean_variance` (learned range, clip_denoised), `normal_kl`,
    `discretized_gaussian_log_likelihood`, `_vb_terms_bpd`, `training_losses` (diffusion.py:17-73, 243-393, 903-1014).
Pinned by tests/golden/make_golden.py::diffusion_case against the REAL modules in train() mode (tests/test_oracle_golden_diffusion.py)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

IN_CH, OUT_CH = 100, 200
N_LATENTS, REF_HEADS = 32, 8
NUM_BUCKETS, MAX_DISTANCE = 32, 64


def default_config(**over):
    cfg = dict(model_channels=512, num_layers=6, in_channels=IN_CH, in_latent_channels=512, out_channels=OUT_CH, num_heads=16)   # diffusion/config.yaml
    cfg.update(over)
    return cfg


def gn_groups(channels):
    """utils.py:124-137"""
    groups = 32
    if channels <= 16:
        groups = 8
    elif channels <= 64:
        groups = 16
    while channels % groups != 0:
        groups = int(groups / 2)
    assert groups > 2
    return groups


# ---------------------------------------------------------------- parameters ----------------------------------------------------------------
def _attn_shapes(s, pre, C, heads):
    s[pre + "norm.weight"] = (C,); s[pre + "norm.bias"] = (C,)
    s[pre + "qkv.weight"] = (3 * C, C, 1); s[pre + "qkv.bias"] = (3 * C,)
    s[pre + "proj_out.weight"] = (C, C, 1); s[pre + "proj_out.bias"] = (C,)
    s[pre + "relative_pos_embeddings.relative_attention_bias.weight"] = (NUM_BUCKETS, heads)


def _res_shapes(s, pre, C):
    s[pre + "in_layers.0.weight"] = (C,); s[pre + "in_layers.0.bias"] = (C,)
    s[pre + "in_layers.2.weight"] = (C, C, 1); s[pre + "in_layers.2.bias"] = (C,)
    s[pre + "emb_layers.1.weight"] = (2 * C, C); s[pre + "emb_layers.1.bias"] = (2 * C,)
    s[pre + "out_layers.0.weight"] = (C,); s[pre + "out_layers.0.bias"] = (C,)
    s[pre + "out_layers.3.weight"] = (C, C, 3); s[pre + "out_layers.3.bias"] = (C,)


def _dl_shapes(s, pre, C, heads):
    _res_shapes(s, pre + "resblk.", C)
    _attn_shapes(s, pre + "attn.", C, heads)


def param_shapes(cfg):
    C, H, L = cfg["model_channels"], cfg["num_heads"], cfg["num_layers"]
    s = {}
    s["inp_block.weight"] = (C, cfg["in_channels"], 3); s["inp_block.bias"] = (C,)
    s["time_embed.0.weight"] = (C, C); s["time_embed.0.bias"] = (C,)
    s["time_embed.2.weight"] = (C, C); s["time_embed.2.bias"] = (C,)
    s["code_norm.weight"] = (C,); s["code_norm.bias"] = (C,)
    s["latent_conditioner.0.weight"] = (C, cfg["in_latent_channels"], 3); s["latent_conditioner.0.bias"] = (C,)
    for i in (1, 2, 3):
        _attn_shapes(s, "latent_conditioner.%d." % i, C, H)
    s["unconditioned_embedding"] = (1, C, 1)
    for i in range(3):
        _dl_shapes(s, "conditioning_timestep_integrator.%d." % i, C, H)
    s["refer_enc.0.weight"] = (C, cfg["in_channels"], 3); s["refer_enc.0.bias"] = (C,)
    for i in (1, 2, 3):
        _attn_shapes(s, "refer_enc.%d." % i, C, H)
    s["refer_enc.4.latents"] = (N_LATENTS, C)
    for c in ("conv_q", "conv_k", "conv_v", "conv_o"):
        s["refer_enc.4.cross_attention.%s.weight" % c] = (C, C, 1); s["refer_enc.4.cross_attention.%s.bias" % c] = (C,)
    s["refer_enc.4.enc.0.weight"] = (C, C, 3); s["refer_enc.4.enc.0.bias"] = (C,)
    for i in (1, 2, 3, 4):
        _attn_shapes(s, "refer_enc.4.enc.%d." % i, C, REF_HEADS)
    s["integrating_conv.weight"] = (C, 2 * C, 1); s["integrating_conv.bias"] = (C,)
    for i in range(L):
        _dl_shapes(s, "layers.%d." % i, C, H)
    for i in range(L, L + 3):
        _res_shapes(s, "layers.%d." % i, C)
    s["out.0.weight"] = (C,); s["out.0.bias"] = (C,)
    s["out.2.weight"] = (cfg["out_channels"], C, 3); s["out.2.bias"] = (cfg["out_channels"],)
    return s


def init_params(cfg, seed=0):
    """numpy-seeded (torch-version independent) values for every tensor; proj_out is NOT left at the reference's zero init, so that the
    attention path carries signal and gradient"""
    rs = np.random.RandomState(seed)
    shapes = param_shapes(cfg)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("norm.weight") or name.endswith("layers.0.weight") or name in ("code_norm.weight", "out.0.weight"):
            v = rs.uniform(0.7, 1.3, size=shp)
        elif name.endswith("bias") and len(shp) == 1:
            v = 0.05 * rs.standard_normal(shp)
        elif "relative_attention_bias" in name:
            v = 0.3 * rs.standard_normal(shp)
        elif name == "unconditioned_embedding":
            v = rs.standard_normal(shp)
        elif name.endswith("latents"):
            v = 0.5 * rs.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[1:]))
            v = rs.standard_normal(shp) * (0.8 / math.sqrt(fan_in))
        out[name] = torch.tensor(v.astype(np.float32))
    return out


# ---------------------------------------------------------------- model ----------------------------------------------------------------
def rel_pos_bucket(Ti, Tj):
    """RelativePositionBias._relative_position_bucket(k_pos - q_pos, causal=False, 32, 64) as an [Ti, Tj] int64 table (xtransformers.py:155-176)"""
    rel = torch.arange(Tj)[None, :] - torch.arange(Ti)[:, None]
    n = -rel
    nb = NUM_BUCKETS // 2
    ret = (n < 0).long() * nb
    n = n.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(MAX_DISTANCE / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


def qkv_attention(qkv, heads, bias_table):
    """QKVAttentionLegacy.forward (utils.py:148-175) with rel_pos = RelativePositionBias(scale = sqrt(ch)); qkv [B, 3C, T] -> [B, C, T]"""
    B, W, T = qkv.shape
    ch = W // (3 * heads)
    q, k, v = qkv.reshape(B * heads, ch * 3, T).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    bias = bias_table[rel_pos_bucket(T, T)]                        # [T, T, H]
    w = (w.reshape(B, heads, T, T) + bias.permute(2, 0, 1)[None] * (ch ** 0.5)).reshape(B * heads, T, T)
    w = torch.softmax(w, dim=-1)
    return torch.einsum("bts,bcs->bct", w, v).reshape(B, -1, T)


def group_norm(x, P, pre):
    return F.group_norm(x, gn_groups(x.shape[1]), P[pre + "weight"], P[pre + "bias"], eps=1e-5)


def attention_block(P, pre, x, heads):
    """AttentionBlock.forward (utils.py:209-215)"""
    qkv = F.conv1d(group_norm(x, P, pre + "norm."), P[pre + "qkv.weight"], P[pre + "qkv.bias"])
    h = qkv_attention(qkv, heads, P[pre + "relative_pos_embeddings.relative_attention_bias.weight"])
    return x + F.conv1d(h, P[pre + "proj_out.weight"], P[pre + "proj_out.bias"])


def resblock(P, pre, x, emb):
    """ResBlock.forward, use_scale_shift_norm, identity skip (aa_model.py:120-135)"""
    h = F.conv1d(F.silu(group_norm(x, P, pre + "in_layers.0.")), P[pre + "in_layers.2.weight"], P[pre + "in_layers.2.bias"])
    eo = F.linear(F.silu(emb), P[pre + "emb_layers.1.weight"], P[pre + "emb_layers.1.bias"])[..., None]
    scale, shift = torch.chunk(eo, 2, dim=1)
    h = group_norm(h, P, pre + "out_layers.0.") * (1 + scale) + shift
    h = F.conv1d(F.silu(h), P[pre + "out_layers.3.weight"], P[pre + "out_layers.3.bias"], padding=1)
    return x + h


def diffusion_layer(P, pre, x, emb, heads):
    return attention_block(P, pre + "attn.", resblock(P, pre + "resblk.", x, emb), heads)


def cross_attention(P, pre, x, c, heads):
    """vc_utils.MultiHeadAttention.forward without window / mask (vc_utils.py:568-600)"""
    q = F.conv1d(x, P[pre + "conv_q.weight"], P[pre + "conv_q.bias"])
    k = F.conv1d(c, P[pre + "conv_k.weight"], P[pre + "conv_k.bias"])
    v = F.conv1d(c, P[pre + "conv_v.weight"], P[pre + "conv_v.bias"])
    B, C, Tq = q.shape
    Tk, dk = k.shape[2], C // heads
    qh = q.view(B, heads, dk, Tq).transpose(2, 3) / math.sqrt(dk)
    kh = k.view(B, heads, dk, Tk).transpose(2, 3)
    vh = v.view(B, heads, dk, Tk).transpose(2, 3)
    p = torch.softmax(qh @ kh.transpose(-2, -1), dim=-1)
    o = (p @ vh).transpose(2, 3).contiguous().view(B, C, Tq)
    return F.conv1d(o, P[pre + "conv_o.weight"], P[pre + "conv_o.bias"])


def ref_encoder(P, pre, x):
    """RefEncoder.forward (aa_model.py:168-177): the channel slice [:, :ref_dim] keeps every channel, the mean runs over latents AND frames"""
    B = x.shape[0]
    lat = P[pre + "latents"].t()[None].expand(B, -1, -1)
    lat = cross_attention(P, pre + "cross_attention.", lat, x, REF_HEADS)
    h = torch.cat((lat, x), -1)
    h = F.conv1d(h, P[pre + "enc.0.weight"], P[pre + "enc.0.bias"], padding=1)
    for i in (1, 2, 3, 4):
        h = attention_block(P, pre + "enc.%d." % i, h, REF_HEADS)
    return h.mean(-1)


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def nearest_index(T_in, T_out):
    """source index of F.interpolate(mode="nearest"): floor(dst * T_in / T_out) computed in fp32 like ATen"""
    scale = np.float32(T_in) / np.float32(T_out)
    return torch.tensor(np.minimum(np.floor(np.arange(T_out, dtype=np.float32) * scale).astype(np.int64), T_in - 1))


def model_forward(P, cfg, x, t, latent, refer, uncond=None, dropped=()):
    """AA_diffusion.forward in train() mode (aa_model.py:256-287).  uncond: bool [B] (samples whose conditioning is replaced) or None;
    dropped: indices of `layers` that are skipped."""
    C, H, L = cfg["model_channels"], cfg["num_heads"], cfg["num_layers"]
    h = F.conv1d(latent, P["latent_conditioner.0.weight"], P["latent_conditioner.0.bias"], padding=1)
    for i in (1, 2, 3):
        h = attention_block(P, "latent_conditioner.%d." % i, h, H)
    r = F.conv1d(refer, P["refer_enc.0.weight"], P["refer_enc.0.bias"], padding=1)
    for i in (1, 2, 3):
        r = attention_block(P, "refer_enc.%d." % i, r, H)
    r = ref_encoder(P, "refer_enc.4.", r)
    le = group_norm(h, P, "code_norm.") + r[..., None]
    if uncond is not None:
        le = torch.where(uncond[:, None, None], P["unconditioned_embedding"].expand(le.shape[0], -1, 1), le)
    le = le[:, :, nearest_index(le.shape[-1], x.shape[-1])]
    te = F.linear(F.silu(F.linear(timestep_embedding(t, C), P["time_embed.0.weight"], P["time_embed.0.bias"])), P["time_embed.2.weight"], P["time_embed.2.bias"])
    for i in range(3):
        le = diffusion_layer(P, "conditioning_timestep_integrator.%d." % i, le, te, H)
    x = F.conv1d(x, P["inp_block.weight"], P["inp_block.bias"], padding=1)
    x = F.conv1d(torch.cat([x, le], dim=1), P["integrating_conv.weight"], P["integrating_conv.bias"])
    for i in range(L + 3):
        if i in dropped:
            assert 0 < i < L + 2
            continue
        x = diffusion_layer(P, "layers.%d." % i, x, te, H) if i < L else resblock(P, "layers.%d." % i, x, te)
    return F.conv1d(F.silu(group_norm(x, P, "out.0.")), P["out.2.weight"], P["out.2.bias"], padding=1)


# ---------------------------------------------------------------- loss ----------------------------------------------------------------
def schedule(n=1000):
    """get_named_beta_schedule('linear', 1000) and the constants GaussianDiffusion.__init__ derives from it (diffusion.py:79-96, 196-229),
    float64 tables like the reference's numpy arrays"""
    betas = np.linspace(1000 / n * 0.0001, 1000 / n * 0.02, n, dtype=np.float64)
    ac = np.cumprod(1.0 - betas)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        sqrt_ac=np.sqrt(ac), sqrt_1mac=np.sqrt(1.0 - ac), sqrt_recip_ac=np.sqrt(1.0 / ac), sqrt_recipm1_ac=np.sqrt(1.0 / ac - 1),
        post_logvar=np.log(np.append(post_var[1], post_var[1:])), log_betas=np.log(betas),
        coef1=betas * np.sqrt(ac_prev) / (1.0 - ac), coef2=(1.0 - ac_prev) * np.sqrt(1.0 - betas) / (1.0 - ac))


def coef_table(t):
    """the eight per-sample coefficients of the loss as an fp32 [B, 8] tensor (the reference's `_extract_into_tensor(...).float()`):
    sqrt_ac, sqrt_1mac, sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2, min_log (posterior_log_variance_clipped), max_log (log beta)"""
    S = schedule()
    tn = np.asarray(t)
    return torch.tensor(np.stack([S[k][tn] for k in ("sqrt_ac", "sqrt_1mac", "sqrt_recip_ac", "sqrt_recipm1_ac", "coef1", "coef2", "post_logvar", "log_betas")],
                                 axis=1).astype(np.float32))


def q_sample(x_start, t, noise):
    c = coef_table(t)
    return c[:, 0, None, None] * x_start + c[:, 1, None, None] * noise


def _cdf(x):
    return 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def loss_terms(model_out, x_start, x_t, noise, t):
    """training_losses' terms for model output [B, 2C, T]: returns (mse [B], vb [B]); loss = (mse + vb).mean() (train.py:172-180)"""
    Cn = x_start.shape[1]
    c = coef_table(t)
    co = lambda i: c[:, i, None, None]
    eps, v = model_out[:, :Cn], model_out[:, Cn:]
    mse = ((noise - eps) ** 2).mean(dim=(1, 2))
    # _vb_terms_bpd with the mean detached (diffusion.py:980-987, 903-928)
    eps_d = eps.detach()
    true_mean = co(4) * x_start + co(5) * x_t
    frac = (v + 1) / 2
    logvar = frac * co(7) + (1 - frac) * co(6)
    pred_x0 = (co(2) * x_t - co(3) * eps_d).clamp(-1, 1)
    mean = co(4) * pred_x0 + co(5) * x_t
    kl = 0.5 * (-1.0 + logvar - co(6) + torch.exp(co(6) - logvar) + ((true_mean - mean) ** 2) * torch.exp(-logvar))
    kl = kl.mean(dim=(1, 2)) / math.log(2.0)
    cx = x_start - mean
    inv_std = torch.exp(-0.5 * logvar)
    cdf_plus, cdf_min = _cdf(inv_std * (cx + 1.0 / 255.0)), _cdf(inv_std * (cx - 1.0 / 255.0))
    log_probs = torch.where(x_start < -0.999, torch.log(cdf_plus.clamp(min=1e-12)),
                            torch.where(x_start > 0.999, torch.log((1.0 - cdf_min).clamp(min=1e-12)), torch.log((cdf_plus - cdf_min).clamp(min=1e-12))))
    nll = -log_probs.mean(dim=(1, 2)) / math.log(2.0)
    vb = torch.where(torch.as_tensor(np.asarray(t)) == 0, nll, kl)
    return mse, vb


def training_loss(P, cfg, x_start, t, noise, latent, refer, uncond=None, dropped=()):
    """one micro-step of ttts/diffusion/train.py:168-180: returns (loss, model_out)"""
    x_t = q_sample(x_start, t, noise)
    out = model_forward(P, cfg, x_t, torch.as_tensor(np.asarray(t)), latent, refer, uncond, dropped)
    mse, vb = loss_terms(out, x_start, x_t, noise, t)
    return (mse + vb).mean(), out


GOLDEN_CFG = dict(model_channels=128, num_layers=3, in_latent_channels=48, num_heads=4)


def golden_inputs(seed=91, B=3, T=24, TL=6, TR=10):
    """seeded inputs of the golden case: normalised mel x_start, timesteps incl. t = 0 (decoder-NLL branch), noise, GPT latent, reference mel,
    one unconditioned sample, two dropped layers (a DiffusionLayer and a ResBlock)"""
    rs = np.random.RandomState(seed)
    f = lambda *s: torch.tensor(rs.standard_normal(s).astype(np.float32))
    x_start = 0.6 * f(B, IN_CH, T)
    x_start[0, :5, :4] = -1.2; x_start[0, 5:9, :4] = 1.1           # exercises the |x| > 0.999 branches of the discretised likelihood
    t = [0, 517, 999][:B]
    return dict(x_start=x_start, t=t, noise=f(B, IN_CH, T), latent=f(B, GOLDEN_CFG["in_latent_channels"], TL), refer=0.5 * f(B, IN_CH, TR),
                uncond=torch.tensor([False, True, False][:B]), dropped=(1, 4))
