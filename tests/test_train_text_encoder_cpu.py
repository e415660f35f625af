"""CPU: the training graph of the prior encoder (ttts_b200/vqvae/train_text_encoder.py, next scope row) over the torch restatement of the
kernel contract (tests/ref_kernels.py) against the REAL reference TextEncoder (tests/golden/text_encoder.npz)."""
import os
import sys

import numpy as np
import torch

from oracle import text_encoder_oracle as TO
from ttts_b200.vqvae.train_encoder import Var
from ttts_b200.vqvae.train_text_encoder import TextEncoderGraph

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402


def test_text_encoder_graph_reproduces_the_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "text_encoder.npz"))
    y, y_lengths, text, text_lengths, ge = TO.golden_inputs()
    graph = TextEncoderGraph(TorchRefKernels(), TO.init_params(seed=8))
    yv, gv = Var(y), Var(ge)
    yo, stats = graph.forward(yv, y_lengths, text, text_lengths, gv)
    m, logs = stats.v[:, :192], stats.v[:, 192:]
    for got, key in ((yo.v, "yo"), (m, "m"), (logs, "logs")):
        assert np.abs(got.numpy() - z[key]).max() <= 5e-5 * max(1.0, np.abs(z[key]).max()), key
    gR = torch.Generator().manual_seed(62)
    R1, R2 = torch.randn(m.shape, generator=gR), torch.randn(m.shape, generator=gR)
    stats.g = torch.cat([R1, R2], dim=1)
    graph.tape.backward()
    grads = graph.grads()
    names = [str(n) for n in z["names"]]
    assert set(names) == set(grads.keys())
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = grads[k]
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 1e-2 * scale + floor, k
    assert np.linalg.norm(yv.g.numpy() - z["dy"]) <= 1e-4 * np.linalg.norm(z["dy"])
    assert np.linalg.norm(gv.g.numpy() - z["dge"]) <= 1e-4 * np.linalg.norm(z["dge"])
