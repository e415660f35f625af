"""CPU: the training graph of the decoder (ttts_b200/vqvae/train_decoder.py, next scope row) over the torch restatement of the kernel contract
(tests/ref_kernels.py): waveform and all parameter gradients against the REAL reference Generator (tests/golden/decoder.npz)."""
import os
import sys

import numpy as np
import torch

from oracle import decoder_oracle as DO
from ttts_b200.vqvae.train_decoder import DecoderGraph

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402


def test_decoder_graph_reproduces_the_reference(golden_dir):
    dec = np.load(os.path.join(golden_dir, "decoder.npz"))
    P = DO.init_params(seed=9)
    graph = DecoderGraph(TorchRefKernels(), P)
    y = graph.forward(torch.tensor(dec["z"]), torch.tensor(dec["g"]))
    assert y.v.shape == dec["y"].shape
    assert np.linalg.norm(y.v.numpy() - dec["y"]) <= 2e-5 * np.linalg.norm(dec["y"])
    R = torch.randn(y.v.shape, generator=torch.Generator().manual_seed(32))
    assert abs(float((y.v * R).sum()) - float(dec["loss"])) <= 1e-4 * max(1.0, abs(float(dec["loss"])))
    grads = graph.backward(R)
    names = [str(n) for n in dec["names"]]
    assert set(names) == set(grads.keys())
    floor = 1e-6 * float(np.sqrt((dec["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = grads[k]
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(dec["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(dec["proj"][i])) <= 1e-2 * scale + floor, k
    # the latent and the conditioning vector receive gradients too (they feed the encoder side of the full step)
    assert graph.z.g is not None and graph.z.g.shape == graph.z.v.shape and graph.g.g.shape == graph.g.v.shape
    zz = torch.tensor(dec["z"]).requires_grad_(True)
    (DO.generator(P, zz, torch.tensor(dec["g"])) * R).sum().backward()
    assert float((graph.z.g - zz.grad).norm()) <= 1e-4 * float(zz.grad.norm())
