"""The state_dict layout of the reference's `AA_diffusion` (ttts/diffusion/aa_model.py:182-254: names and shapes a checkpoint of the reference
carries) and a fresh initialisation in the reference's style (torch's Conv1d / Linear / GroupNorm / Embedding defaults, zeroed `proj_out`
of every AttentionBlock -- utils.py:203 `zero_module`, latents ~ N(0, 0.02), unconditioned embedding ~ N(0, 1))."""
import math

import torch

N_LATENTS, REF_HEADS, NUM_BUCKETS = 32, 8, 32


def default_config(**over):
    cfg = dict(model_channels=512, num_layers=6, in_channels=100, in_latent_channels=512, out_channels=200, num_heads=16)   # ttts/diffusion/config.yaml
    cfg.update(over)
    return cfg


def param_shapes(cfg):
    C, H, L = cfg["model_channels"], cfg["num_heads"], cfg["num_layers"]
    s = {}

    def attn(pre, heads):
        s[pre + "norm.weight"] = (C,); s[pre + "norm.bias"] = (C,)
        s[pre + "qkv.weight"] = (3 * C, C, 1); s[pre + "qkv.bias"] = (3 * C,)
        s[pre + "proj_out.weight"] = (C, C, 1); s[pre + "proj_out.bias"] = (C,)
        s[pre + "relative_pos_embeddings.relative_attention_bias.weight"] = (NUM_BUCKETS, heads)

    def res(pre):
        s[pre + "in_layers.0.weight"] = (C,); s[pre + "in_layers.0.bias"] = (C,)
        s[pre + "in_layers.2.weight"] = (C, C, 1); s[pre + "in_layers.2.bias"] = (C,)
        s[pre + "emb_layers.1.weight"] = (2 * C, C); s[pre + "emb_layers.1.bias"] = (2 * C,)
        s[pre + "out_layers.0.weight"] = (C,); s[pre + "out_layers.0.bias"] = (C,)
        s[pre + "out_layers.3.weight"] = (C, C, 3); s[pre + "out_layers.3.bias"] = (C,)

    def dl(pre):
        res(pre + "resblk."); attn(pre + "attn.", H)

    s["inp_block.weight"] = (C, cfg["in_channels"], 3); s["inp_block.bias"] = (C,)
    s["time_embed.0.weight"] = (C, C); s["time_embed.0.bias"] = (C,)
    s["time_embed.2.weight"] = (C, C); s["time_embed.2.bias"] = (C,)
    s["code_norm.weight"] = (C,); s["code_norm.bias"] = (C,)
    s["latent_conditioner.0.weight"] = (C, cfg["in_latent_channels"], 3); s["latent_conditioner.0.bias"] = (C,)
    for i in (1, 2, 3):
        attn("latent_conditioner.%d." % i, H)
    s["unconditioned_embedding"] = (1, C, 1)
    for i in range(3):
        dl("conditioning_timestep_integrator.%d." % i)
    s["refer_enc.0.weight"] = (C, cfg["in_channels"], 3); s["refer_enc.0.bias"] = (C,)
    for i in (1, 2, 3):
        attn("refer_enc.%d." % i, H)
    s["refer_enc.4.latents"] = (N_LATENTS, C)
    for c in ("conv_q", "conv_k", "conv_v", "conv_o"):
        s["refer_enc.4.cross_attention.%s.weight" % c] = (C, C, 1); s["refer_enc.4.cross_attention.%s.bias" % c] = (C,)
    s["refer_enc.4.enc.0.weight"] = (C, C, 3); s["refer_enc.4.enc.0.bias"] = (C,)
    for i in (1, 2, 3, 4):
        attn("refer_enc.4.enc.%d." % i, REF_HEADS)
    s["integrating_conv.weight"] = (C, 2 * C, 1); s["integrating_conv.bias"] = (C,)
    for i in range(L):
        dl("layers.%d." % i)
    for i in range(L, L + 3):
        res("layers.%d." % i)
    s["out.0.weight"] = (C,); s["out.0.bias"] = (C,)
    s["out.2.weight"] = (cfg["out_channels"], C, 3); s["out.2.bias"] = (cfg["out_channels"],)
    return s


def init_params(cfg, seed=0, device="cpu", zero_proj_out=True):
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(cfg)
    out = {}
    for name, shp in shapes.items():
        norm_w = name.endswith("norm.weight") or name.endswith("layers.0.weight") or name in ("code_norm.weight", "out.0.weight")
        norm_b = name.endswith("norm.bias") or name.endswith("layers.0.bias") or name in ("code_norm.bias", "out.0.bias")
        if norm_w:
            v = torch.ones(shp)
        elif norm_b:
            v = torch.zeros(shp)
        elif zero_proj_out and "proj_out." in name:
            v = torch.zeros(shp)
        elif "relative_attention_bias" in name or name == "unconditioned_embedding":
            v = torch.randn(shp, generator=g)
        elif name.endswith("latents"):
            v = 0.02 * torch.randn(shp, generator=g)
        else:
            w_name = name[:-4] + "weight" if name.endswith("bias") else name
            fan_in = 1
            for d in shapes[w_name][1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)                         # kaiming_uniform(a = sqrt(5)) and the bias bound of torch's Conv1d / Linear
            v = (torch.rand(shp, generator=g) * 2 - 1) * bound
        out[name] = v.to(device)
    return out
