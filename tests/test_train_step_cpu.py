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


def test_full_step_order_and_update_reproduce_the_real_reference(golden_dir):
    """train_step.TrainStep (synthesis -> D loss -> optim_d -> adversarial losses through the UPDATED discriminators -> optim_g) with the
    trainer's torch AdamW plugged in: losses and the update of every net_g / net_d parameter tensor against ONE real optimisation step of the
    reference (tests/golden/vqvae_full_step.npz).  The generator loss drops from 5.53 to 2.53 when the discriminators are stepped first, so
    the order is visible in the numbers."""
    import make_golden as MG
    from ref_kernels import TorchAdamW
    from ttts_b200.vqvae.train_step import TrainStep
    z = np.load(os.path.join(golden_dir, "vqvae_full_step.npz"))
    G, D = MG.step_params()
    wav, lengths, text, text_lengths, E = MG.step_inputs()
    spec = torch.tensor(V.spectrogram(wav.numpy()))
    torch.manual_seed(0)
    eps_p, eps_q = torch.randn(3, 192, 36), torch.randn(3, 192, 36)
    ids = (torch.rand([3]) * (lengths - 8 + 1)).to(torch.long).tolist()
    ts = TrainStep(TorchRefKernels(), G, D, optimizer=TorchAdamW)
    out = ts.step(wav, spec, lengths, text, text_lengths, E, eps_p, eps_q, ids, 8)
    for key in ("loss_disc", "loss_gen", "loss_fm", "total"):
        assert abs(float(out[key]) - float(z[key])) <= 3e-4 * max(1.0, abs(float(z[key]))), (key, float(out[key]), float(z[key]))
    for tag, opt, before in (("g", ts.opt_g, G), ("d", ts.opt_d, D)):
        names = [str(n) for n in z[tag + "_names"]]
        after = opt.params()
        assert set(names) == set(after.keys())
        bad = []
        for i, k in enumerate(names):
            if k.endswith("conv_k.bias") or k.endswith("w_ks.bias"):
                # attention KEY biases: their gradient is exactly zero in exact arithmetic (softmax is shift-invariant), so what AdamW turns
                # into a +-lr update is fp32 rounding noise -- in the reference as much as here.  Only the size of the update is comparable.
                assert abs(float((after[k] - before[k]).norm()) - float(z[tag + "_norm"][i])) <= 0.5 * float(z[tag + "_norm"][i]) + 1e-9, k
                continue
            dlt = after[k] - before[k]
            d = torch.randn(dlt.shape, generator=torch.Generator().manual_seed(i))
            scale = float(z[tag + "_norm"][i])
            # the first AdamW step is sign-like (|update| = lr wherever |g| >> eps = 1e-9): elements whose gradient sits at fp32 noise level flip
            # freely, 0.1 % of flipped elements already move the projection by 6 % of the norm -> 15 % here, 2 % on the norm
            if not (abs(float(dlt.norm()) - scale) <= 2e-2 * scale + 1e-9 and abs(float((dlt * d).sum()) - float(z[tag + "_proj"][i])) <= 0.15 * scale + 1e-9):
                bad.append((k, float(dlt.norm()), scale, abs(float((dlt * d).sum()) - float(z[tag + "_proj"][i])) / scale))
        assert len(bad) <= len(names) // 200, bad[:10]
