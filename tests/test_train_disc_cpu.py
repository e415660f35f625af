"""CPU: the training graph of the discriminators + adversarial losses (ttts_b200/vqvae/train_disc.py, next scope row) over the torch
restatement of the kernel contract (tests/ref_kernels.py), against the REAL reference MultiPeriodDiscriminator / losses (tests/golden/disc.npz)."""
import os
import sys

import numpy as np
import torch

from oracle import disc_oracle as DO
from ttts_b200.vqvae.train_disc import PERIODS, DiscriminatorGraph
from ttts_b200.vqvae.train_encoder import Var

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402


def _logits_like_reference(out, d, B):
    """our period discriminators keep the p columns in the batch axis: [B p, 1, R] -> the reference's flatten of [B, 1, R, p]"""
    v = out.v
    if d == 0:
        return v.reshape(B, -1)
    p = PERIODS[d - 1]
    return v.view(B, p, -1).permute(0, 2, 1).reshape(B, -1)


def test_discriminator_step(golden_dir):
    z = np.load(os.path.join(golden_dir, "disc.npz"))
    graph = DiscriminatorGraph(TorchRefKernels(), DO.init_params(seed=4))
    y, y_hat = torch.tensor(z["y"]), torch.tensor(z["y_hat"])
    real, _ = graph.forward(y)
    gen, _ = graph.forward(y_hat)
    for d in range(6):
        for out, key in ((real[d], "d%d_real" % d), (gen[d], "d%d_gen" % d)):
            got = _logits_like_reference(out, d, 2).numpy()
            assert got.shape == z[key].shape and np.abs(got - z[key]).max() <= 2e-5 * max(1.0, np.abs(z[key]).max()), key
    loss = graph.discriminator_loss(real, gen)
    assert abs(float(loss.v) - float(z["loss_d"])) <= 1e-5 * float(z["loss_d"])
    grads = graph.backward(loss)
    names = [str(n) for n in z["names"]]
    assert set(names) == set(grads.keys())
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = grads[k]
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 1e-2 * scale + floor, k


def test_generator_step_losses_and_waveform_gradient(golden_dir):
    z = np.load(os.path.join(golden_dir, "disc.npz"))
    graph = DiscriminatorGraph(TorchRefKernels(), DO.init_params(seed=4))
    y_hat = Var(torch.tensor(z["y_hat"]))
    _, fmap_r = graph.forward(torch.tensor(z["y"]))
    gen, fmap_g = graph.forward(y_hat)
    loss_gen, loss_fm = graph.generator_losses(gen, fmap_r, fmap_g)
    assert abs(float(loss_gen.v) - float(z["loss_gen"])) <= 1e-5 * float(z["loss_gen"])
    assert abs(float(loss_fm.v) - float(z["loss_fm"])) <= 1e-5 * float(z["loss_fm"])
    graph.backward(graph.ops.add(loss_gen, loss_fm))
    want = z["dy_hat"]
    assert np.linalg.norm(y_hat.g.numpy() - want) <= 1e-4 * np.linalg.norm(want)
