"""CPU ORACLE (test infrastructure, never the product path) for the NEXT scope row: the normalising flow of SynthesizerTrn,
`ResidualCouplingBlock(192, 192, 5, 1, 4, gin_channels=512)` (ttts/vqvae/vq2.py:209-246, built at :829-831): four mean-only
`ResidualCouplingLayer`s (modules.py:405-459: pre 1x1 -> 4-layer WN conditioned on g -> post 1x1 -> x1 = m + x1 * mask) each followed by a
channel flip, and the KL term it feeds (losses.py:47-61, restated in disc_oracle.kl_loss).  Pinned by tests/golden/make_golden.py::flow_case
against the REAL reference module (tests/test_oracle_golden_flow.py)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

CH, HID, GIN, KS, NL, NF = 192, 192, 512, 5, 4, 4


def param_shapes():
    s = {}
    for f in range(NF):
        p = "flows.%d." % (2 * f)
        s[p + "pre.weight"] = (HID, CH // 2, 1); s[p + "pre.bias"] = (HID,)
        s[p + "enc.cond_layer.weight_g"] = (2 * HID * NL, 1, 1); s[p + "enc.cond_layer.weight_v"] = (2 * HID * NL, GIN, 1)
        s[p + "enc.cond_layer.bias"] = (2 * HID * NL,)
        for i in range(NL):
            s[p + "enc.in_layers.%d.weight_g" % i] = (2 * HID, 1, 1); s[p + "enc.in_layers.%d.weight_v" % i] = (2 * HID, HID, KS)
            s[p + "enc.in_layers.%d.bias" % i] = (2 * HID,)
            co = 2 * HID if i < NL - 1 else HID
            s[p + "enc.res_skip_layers.%d.weight_g" % i] = (co, 1, 1); s[p + "enc.res_skip_layers.%d.weight_v" % i] = (co, HID, 1)
            s[p + "enc.res_skip_layers.%d.bias" % i] = (co,)
        s[p + "post.weight"] = (CH // 2, HID, 1); s[p + "post.bias"] = (CH // 2,)
    return s


def init_params(seed=0):
    """numpy-seeded; `post` is NOT zero here (the reference zero-initialises it, which would make the flow the identity and the test vacuous)"""
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
    for name in list(out):
        if name.endswith("weight_g"):
            vn = out[name[:-1] + "v"].flatten(1).norm(dim=1).view(out[name].shape)
            out[name] = out[name] * vn * 0.9
    return out


def _wn(P, prefix):
    g, v = P[prefix + "weight_g"], P[prefix + "weight_v"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def wn(P, pre, x, x_mask, g):
    """modules.py:136-221 with n_layers = 4"""
    out = torch.zeros_like(x)
    gc = F.conv1d(g, _wn(P, pre + "cond_layer."), P[pre + "cond_layer.bias"])
    for i in range(NL):
        x_in = F.conv1d(x, _wn(P, pre + "in_layers.%d." % i), P[pre + "in_layers.%d.bias" % i], padding=(KS - 1) // 2)
        a = x_in + gc[:, i * 2 * HID:(i + 1) * 2 * HID, :]
        acts = torch.tanh(a[:, :HID]) * torch.sigmoid(a[:, HID:])
        rs = F.conv1d(acts, _wn(P, pre + "res_skip_layers.%d." % i), P[pre + "res_skip_layers.%d.bias" % i])
        if i < NL - 1:
            x = (x + rs[:, :HID]) * x_mask
            out = out + rs[:, HID:]
        else:
            out = out + rs
    return out * x_mask


def flow(P, x, x_mask, g):
    """ResidualCouplingBlock.forward, reverse=False (vq2.py:238-241)"""
    for f in range(NF):
        p = "flows.%d." % (2 * f)
        x0, x1 = torch.split(x, [CH // 2] * 2, 1)
        h = F.conv1d(x0, P[p + "pre.weight"], P[p + "pre.bias"]) * x_mask
        h = wn(P, p + "enc.", h, x_mask, g)
        m = F.conv1d(h, P[p + "post.weight"], P[p + "post.bias"]) * x_mask
        x = torch.cat([x0, m + x1 * x_mask], 1)
        x = torch.flip(x, [1])
    return x


def golden_inputs():
    """the seeded inputs of tests/golden/make_golden.py::flow_case (regenerated, not stored): z, g, mask [B,1,T], logs_q, m_p, logs_p"""
    g0 = torch.Generator().manual_seed(51)
    z = torch.randn(3, 192, 36, generator=g0)
    ge = torch.randn(3, 512, 1, generator=g0)
    lengths = torch.tensor([36, 29, 11])
    mask = (torch.arange(36)[None, :] < lengths[:, None]).float().unsqueeze(1)
    logs_q, m_p, logs_p = [0.3 * torch.randn(3, 192, 36, generator=g0) for _ in range(3)]
    return z, ge, mask, logs_q, m_p, logs_p
