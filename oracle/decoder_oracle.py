"""CPU ORACLE (test infrastructure, never the product path) for the NEXT scope row: the VQ-VAE decoder, i.e. the
HiFi-GAN-style `Generator` of SynthesizerTrn (ttts/vqvae/vq2.py:341-416, built at :798-807 with the hyper-parameters of
ttts/vqvae/config.json:65-92): conv_pre(192 -> 512, k7) + cond(g), five [leaky_relu(0.1) -> weight-normed ConvTranspose1d (rates 10, 8, 2, 2,
2; kernels 16, 16, 8, 2, 2) -> mean of three ResBlock1 (kernels 3 / 7 / 11, dilations 1 / 3 / 5)], leaky_relu(0.01), conv_post(16 -> 1, k7,
no bias), tanh.  Plain torch functional ops on a name -> tensor dict with the reference's state_dict names (`dec.` prefix dropped).

No kernels exist for this row yet; the oracle is pinned first, as the scope order demands: tests/golden/make_golden.py::decoder_case runs the
REAL reference Generator on CPU and tests/test_oracle_golden_decoder.py compares outputs and gradients.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .encoder_oracle import resblock1

INTER, GIN, UP_INIT = 192, 512, 512
RATES = [10, 8, 2, 2, 2]
KSZ = [16, 16, 8, 2, 2]
RES_K = (3, 7, 11)


def param_shapes():
    s = {"conv_pre.weight": (UP_INIT, INTER, 7), "conv_pre.bias": (UP_INIT,), "cond.weight": (UP_INIT, GIN, 1), "cond.bias": (UP_INIT,)}
    for i, k in enumerate(KSZ):
        cin, cout = UP_INIT // (2 ** i), UP_INIT // (2 ** (i + 1))
        s["ups.%d.weight_g" % i] = (cin, 1, 1)                  # old-style weight_norm over dim 0 of the [Cin, Cout, K] transposed-conv weight
        s["ups.%d.weight_v" % i] = (cin, cout, k)
        s["ups.%d.bias" % i] = (cout,)
        for j, rk in enumerate(RES_K):
            for cs in ("convs1", "convs2"):
                for t in range(3):
                    p = "resblocks.%d.%s.%d." % (i * 3 + j, cs, t)
                    s[p + "parametrizations.weight.original0"] = (cout, 1, 1)
                    s[p + "parametrizations.weight.original1"] = (cout, cout, rk)
                    s[p + "bias"] = (cout,)
    s["conv_post.weight"] = (1, UP_INIT // 32, 7)
    return s


def init_params(seed=0):
    """Deterministic (numpy-seeded) parameters, same recipe as encoder_oracle.init_params."""
    rs = np.random.RandomState(seed)
    out = {}
    shapes = param_shapes()
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("weight_g") or name.endswith("original0"):
            v = rs.uniform(0.6, 1.4, size=shp)
        elif name.endswith("bias"):
            v = 0.05 * rs.standard_normal(shp)
        else:
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            v = rs.standard_normal(shp) / math.sqrt(fan_in)
        out[name] = torch.tensor(v, dtype=torch.float32)
    for name in list(out):
        if name.endswith("weight_g") or name.endswith("original0"):
            vname = name[:-len("weight_g")] + "weight_v" if name.endswith("weight_g") else name[:-1] + "1"
            vn = out[vname].flatten(1).norm(dim=1).view(out[name].shape)
            out[name] = out[name] * vn * 0.9
    return out


def generator(P, x, g=None):
    """x [B, 192, T], g [B, 512, 1] or None -> waveform [B, 1, 640 T]   (vq2.py:389-408)"""
    x = F.conv1d(x, P["conv_pre.weight"], P["conv_pre.bias"], padding=3)
    if g is not None:
        x = x + F.conv1d(g, P["cond.weight"], P["cond.bias"])
    for i, (u, k) in enumerate(zip(RATES, KSZ)):
        x = F.leaky_relu(x, 0.1)
        gw, v = P["ups.%d.weight_g" % i], P["ups.%d.weight_v" % i]
        w = gw * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)
        x = F.conv_transpose1d(x, w, P["ups.%d.bias" % i], stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RES_K):
            r = resblock1(P, "resblocks.%d." % (i * 3 + j), x, rk)
            xs = r if xs is None else xs + r
        x = xs / len(RES_K)
    x = F.leaky_relu(x)                                      # default slope 0.01 here, as in the reference (vq2.py:404)
    x = F.conv1d(x, P["conv_post.weight"], None, padding=3)
    return torch.tanh(x)
