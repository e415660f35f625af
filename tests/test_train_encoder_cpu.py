"""CPU: the training graph of the encode half (ttts_b200/vqvae/train_encoder.py: tape + op set, next scope row) run over the torch restatement
of its kernel contract (tests/ref_kernels.py).  All 414 parameter gradients must equal those of the REAL reference modules
(tests/golden/encoder_grads.npz, minted by make_golden.py::encoder_case) -- this pins the graph's wiring and the tape; the CUDA kernels behind
the same contract are checked separately (tests/test_emu_*.py on the CPU emulation, tests/test_gpu_encoder.py on hardware)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as EO
from oracle import vq_mel_oracle as V
from ttts_b200.vqvae.train_encoder import EncoderGraph

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402


@pytest.fixture(scope="module")
def enc(golden_dir):
    return np.load(os.path.join(golden_dir, "encoder.npz"))


def test_training_graph_reproduces_the_reference_gradients(enc, golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_grads.npz"))
    P = EO.init_params(seed=5)
    wav = torch.tensor(enc["wav"])
    spec = torch.tensor(V.spectrogram(enc["wav"]))
    graph = EncoderGraph(TorchRefKernels(), P)
    z, x = graph.forward(spec, wav, lengths=torch.tensor(enc["lengths"]), eps=torch.tensor(enc["eps"]))
    # forward values against the reference's own outputs
    assert np.abs(z.v.numpy() - enc["z"]).max() <= 5e-4 * np.abs(enc["z"]).max()
    assert np.abs(x.v.numpy() - enc["x"]).max() <= 5e-4 * np.abs(enc["x"]).max()
    R = torch.randn(3, 192, 36, generator=torch.Generator().manual_seed(123))
    loss = float((z.v * R).sum() + 0.5 * (x.v ** 2).mean())
    assert abs(loss - float(g["loss"])) < 2e-3 * abs(float(g["loss"]))
    grads = graph.backward(dz=R, dx=x.v / x.v.numel())                       # L = <z, R> + 0.5 mean(x^2)
    names = [str(n) for n in g["names"]]
    assert set(names) == set(grads.keys())
    floor = 1e-6 * float(np.sqrt((g["norm"] ** 2).sum()))
    num = den = 0.0
    for i, k in enumerate(names):
        gk = grads[k]
        assert tuple(gk.shape) == tuple(P[k].shape), k
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(g["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(g["proj"][i])) <= 1e-2 * scale + floor, k
        num += (float(gk.norm()) - scale) ** 2
        den += scale ** 2
    assert (num / den) ** 0.5 < 1e-3
    for key in g.files:
        if key.startswith("gradfull/"):
            k = key[len("gradfull/"):]
            assert np.linalg.norm(grads[k].numpy() - g[key]) <= 2e-3 * np.linalg.norm(g[key]) + floor, k
    assert graph.tape.steps == []                                             # the tape is consumed


def test_product_backend_refuses_to_run_off_gpu():
    from ttts_b200.vqvae import train_encoder as TE
    if hasattr(TE, "CudaKernels") and not torch.cuda.is_available():
        with pytest.raises(Exception):
            TE.CudaKernels().add(torch.zeros(2), torch.zeros(2))
