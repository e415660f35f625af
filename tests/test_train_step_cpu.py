"""CPU: the generator half of one VQ-VAE-GAN train step assembled on ONE tape (ttts_b200/vqvae/train_step.py, next scope row) over the torch
restatement of the kernel contract (tests/ref_kernels.py), against the REAL reference step (tests/golden/vqvae_step.npz, minted by
make_golden.py::vqvae_step_case from SynthesizerTrn + MultiPeriodDiscriminator + the trainer's loss formulas): the five losses and the
gradient of loss_gen_all with respect to all 1455 net_g parameter tensors."""
import os
import sys

import numpy as np
import torch

from oracle import vq_mel_oracle as V
from ttts_b200.vqvae.train_step import GeneratorStep

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from ref_kernels import TorchRefKernels  # noqa: E402


def test_generator_step_reproduces_the_real_reference(golden_dir):
    import make_golden as MG                                   # only its parameter / input assemblers (no reference import happens here)
    z = np.load(os.path.join(golden_dir, "vqvae_step.npz"))
    G, D = MG.step_params()
    wav, lengths, text, text_lengths, E = MG.step_inputs()
    spec = torch.tensor(V.spectrogram(wav.numpy()))
    # the reference's draws, in its order: enc_p's randn_like, enc_q's randn_like, rand_slice_segments' rand (vq2.py:744 twice, commons.py:62)
    torch.manual_seed(0)
    eps_p, eps_q = torch.randn(3, 192, 36), torch.randn(3, 192, 36)
    ids = (torch.rand([3]) * (lengths - 8 + 1)).to(torch.long)
    assert ids.tolist() == z["ids_slice"].tolist()
    step = GeneratorStep(TorchRefKernels(), G, D)
    out = step.forward(wav, spec, lengths, text, text_lengths, E, eps_p, eps_q, ids.tolist(), 8)
    assert abs(float(out["y_hat"].v.sum()) - float(z["y_hat_sum"])) <= 1e-3 * max(1.0, abs(float(z["y_hat_sum"])))
    for key in ("loss_gen", "loss_fm", "loss_mel", "kl_ssl", "loss_kl", "total"):
        assert abs(float(out[key].v) - float(z[key])) <= 2e-4 * max(1.0, abs(float(z[key]))), (key, float(out[key].v), float(z[key]))
    grads = step.backward()
    names = [str(n) for n in z["names"]]
    assert set(names) == set(grads.keys()), sorted(set(names) ^ set(grads.keys()))[:8]
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    bad = []
    for i, k in enumerate(names):
        gk = grads[k]
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        if not (abs(float(gk.norm()) - scale) <= 3e-3 * scale + floor and abs(float((gk * d).sum()) - float(z["proj"][i])) <= 1.5e-2 * scale + floor):
            bad.append((k, float(gk.norm()), scale))
    assert not bad, bad[:10]
