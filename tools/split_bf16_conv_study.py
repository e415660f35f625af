"""Numerical study for the round-2 plan: what happens to the encoder's outputs and to the VQ codes if the ResBlock /
WN convolutions run on bf16 tensor cores with split operands (x = hi + lo, w = hi + lo; products hi*hi + hi*lo + lo*hi, fp32 accumulation)
instead of fp32 FMAs?  Pure CPU: patches torch.nn.functional.conv1d inside the encoder oracle and compares against the golden vectors
minted from the real reference (tests/golden/encoder.npz).  Variants: 'bf16' (plain bf16 operands), 'split3' (the plan), 'split4' (+ lo*lo).

    python tools/split_bf16_conv_study.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

from oracle import encoder_oracle as EO
from oracle import vq_mel_oracle as V

_conv = F.conv1d


def split(t):
    hi = t.to(torch.bfloat16).float()
    lo = (t - hi).to(torch.bfloat16).float()
    return hi, lo


def make_conv(mode, min_cin=16):
    def conv(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        if mode == "fp32" or groups != 1 or w.shape[1] < min_cin:
            return _conv(x, w, b, stride, padding, dilation, groups)
        kw = dict(stride=stride, padding=padding, dilation=dilation)
        xh, xl = split(x)
        wh, wl = split(w)
        if mode == "bf16":
            y = _conv(xh, wh, None, **kw)
        else:
            y = _conv(xh, wh, None, **kw) + _conv(xh, wl, None, **kw) + _conv(xl, wh, None, **kw)
            if mode == "split4":
                y = y + _conv(xl, wl, None, **kw)
        return y if b is None else y + b.view(1, -1, 1)
    return conv


def rel(a, b):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def main():
    enc = np.load(os.path.join(ROOT, "tests", "golden", "encoder.npz"))
    P = EO.init_params(seed=5)
    wav = torch.tensor(enc["wav"])
    spec = torch.tensor(V.spectrogram(enc["wav"]))
    lengths = torch.tensor(enc["lengths"])
    eps = torch.tensor(enc["eps"])
    xn = np.ascontiguousarray(enc["x"].transpose(0, 2, 1)).reshape(-1, 192)
    margin = V.vq_margin(xn, enc["E"], enc["codes"].reshape(-1))
    print("%-7s %10s %10s %10s %10s   %s" % ("convs", "rel(m)", "rel(logs)", "rel(z)", "rel(x)", "code flips (of %d), largest margin of a flipped code" % xn.shape[0]))
    for mode in ("fp32", "bf16", "split3", "split4"):
        EO.F.conv1d = make_conv(mode)
        try:
            with torch.no_grad():
                out = EO.encode(P, spec, wav, lengths=lengths, eps=eps, codebook=enc["E"])
        finally:
            EO.F.conv1d = _conv
        codes = np.asarray(out["codes"]).reshape(-1)
        flips = codes != enc["codes"].reshape(-1)
        print("%-7s %10.2e %10.2e %10.2e %10.2e   %d, %.2e" % (mode, rel(out["m"], enc["m"]), rel(out["logs"], enc["logs"]), rel(out["z"], enc["z"]),
                                                               rel(out["x"], enc["x"]), int(flips.sum()), float(margin[flips].max()) if flips.any() else 0.0))


if __name__ == "__main__":
    main()
