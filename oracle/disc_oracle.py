"""CPU ORACLE (test infrastructure, never the product path) for the NEXT scope row: the adversarial half of the VQ-VAE-GAN
train step -- `MultiPeriodDiscriminator` = one scale discriminator (grouped Conv1d stack) + five period discriminators ((k,1) Conv2d over
the waveform folded to [T/p, p]) (ttts/vqvae/vq2.py:418-551) and the losses the trainer forms from them (ttts/vqvae/losses.py:7-61:
feature_loss, discriminator_loss, generator_loss, kl_loss).  Plain torch functional ops on a name -> tensor dict with the reference's
state_dict names (old-style weight_norm: `weight_g` / `weight_v`).

No kernels exist for this row yet; the oracle is pinned first, as the scope order demands: tests/golden/make_golden.py::disc_case runs the
REAL reference modules on CPU and tests/test_oracle_golden_disc.py compares outputs, losses and gradients.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PERIODS = [2, 3, 5, 7, 11]
LRELU_SLOPE = 0.1
# DiscriminatorS: (cin, cout, kernel, stride, groups, padding)                                     vq2.py:498-507
S_CONVS = [(1, 16, 15, 1, 1, 7), (16, 64, 41, 4, 4, 20), (64, 256, 41, 4, 16, 20), (256, 1024, 41, 4, 64, 20), (1024, 1024, 41, 4, 256, 20),
           (1024, 1024, 5, 1, 1, 2)]
# DiscriminatorP: (cin, cout, stride) with kernel (5,1), padding (2,0)                              vq2.py:425-470
P_CONVS = [(1, 32, 3), (32, 128, 3), (128, 512, 3), (512, 1024, 3), (1024, 1024, 1)]


def param_shapes():
    s = {}
    p = "discriminators.0."
    for i, (cin, cout, k, st, g, pad) in enumerate(S_CONVS):
        s[p + "convs.%d.weight_g" % i] = (cout, 1, 1)
        s[p + "convs.%d.weight_v" % i] = (cout, cin // g, k)
        s[p + "convs.%d.bias" % i] = (cout,)
    s[p + "conv_post.weight_g"] = (1, 1, 1)
    s[p + "conv_post.weight_v"] = (1, 1024, 3)
    s[p + "conv_post.bias"] = (1,)
    for d in range(len(PERIODS)):
        p = "discriminators.%d." % (d + 1)
        for i, (cin, cout, st) in enumerate(P_CONVS):
            s[p + "convs.%d.weight_g" % i] = (cout, 1, 1, 1)
            s[p + "convs.%d.weight_v" % i] = (cout, cin, 5, 1)
            s[p + "convs.%d.bias" % i] = (cout,)
        s[p + "conv_post.weight_g"] = (1, 1, 1, 1)
        s[p + "conv_post.weight_v"] = (1, 1024, 3, 1)
        s[p + "conv_post.bias"] = (1,)
    return s


def init_params(seed=0):
    """Deterministic (numpy-seeded) parameters, same recipe as encoder_oracle / decoder_oracle.init_params."""
    rs = np.random.RandomState(seed)
    out = {}
    shapes = param_shapes()
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("weight_g"):
            v = rs.uniform(0.6, 1.4, size=shp)
        elif name.endswith("bias"):
            v = 0.05 * rs.standard_normal(shp)
        else:
            v = rs.standard_normal(shp) / math.sqrt(int(np.prod(shp[1:])))
        out[name] = torch.tensor(v, dtype=torch.float32)
    return out


def _wn(P, prefix):
    """old-style torch.nn.utils.weight_norm (dim 0): w = g * v / ||v|| per output channel"""
    v, g = P[prefix + "weight_v"], P[prefix + "weight_g"]
    return g * v / v.flatten(1).norm(dim=1).view(g.shape)


def disc_s(P, x, prefix="discriminators.0."):
    """DiscriminatorS.forward (vq2.py:511-522): x [B, 1, T] -> (flat logits [B, T'], feature maps)"""
    fmap = []
    for i, (cin, cout, k, st, g, pad) in enumerate(S_CONVS):
        x = F.conv1d(x, _wn(P, prefix + "convs.%d." % i), P[prefix + "convs.%d.bias" % i], stride=st, padding=pad, groups=g)
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmap.append(x)
    x = F.conv1d(x, _wn(P, prefix + "conv_post."), P[prefix + "conv_post.bias"], padding=1)
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def disc_p(P, x, period, prefix):
    """DiscriminatorP.forward (vq2.py:473-493): reflect-pad T to a multiple of the period, fold to [B, 1, T/p, p], (5,1) convolutions"""
    fmap = []
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    for i, (cin, cout, st) in enumerate(P_CONVS):
        x = F.conv2d(x, _wn(P, prefix + "convs.%d." % i), P[prefix + "convs.%d.bias" % i], stride=(st, 1), padding=(2, 0))
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmap.append(x)
    x = F.conv2d(x, _wn(P, prefix + "conv_post."), P[prefix + "conv_post.bias"], padding=(1, 0))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def mpd(P, y, y_hat):
    """MultiPeriodDiscriminator.forward (vq2.py:536-551)"""
    y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
    for d in range(1 + len(PERIODS)):
        if d == 0:
            r, fr = disc_s(P, y)
            g, fg = disc_s(P, y_hat)
        else:
            pre = "discriminators.%d." % d
            r, fr = disc_p(P, y, PERIODS[d - 1], pre)
            g, fg = disc_p(P, y_hat, PERIODS[d - 1], pre)
        y_d_rs.append(r); y_d_gs.append(g); fmap_rs.append(fr); fmap_gs.append(fg)
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs


def feature_loss(fmap_r, fmap_g):
    """losses.py:7-15"""
    loss = 0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            loss = loss + torch.mean(torch.abs(rl.detach() - gl))
    return loss * 2


def discriminator_loss(disc_real, disc_gen):
    """losses.py:18-32 (least-squares GAN)"""
    loss = 0
    for dr, dg in zip(disc_real, disc_gen):
        loss = loss + torch.mean((1 - dr) ** 2) + torch.mean(dg ** 2)
    return loss


def generator_loss(disc_gen):
    """losses.py:35-44"""
    loss = 0
    for dg in disc_gen:
        loss = loss + torch.mean((1 - dg) ** 2)
    return loss


def kl_loss(z_p, logs_q, m_p, logs_p, z_mask):
    """losses.py:47-61"""
    kl = logs_p - logs_q - 0.5
    kl = kl + 0.5 * ((z_p - m_p) ** 2) * torch.exp(-2.0 * logs_p)
    return torch.sum(kl * z_mask) / torch.sum(z_mask)
