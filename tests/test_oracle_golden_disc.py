"""CPU: the discriminator / GAN-loss oracle (oracle/disc_oracle.py: MultiPeriodDiscriminator + losses of the VQ-VAE-GAN step, the next scope
row) against golden vectors minted from the REAL reference modules (tests/golden/make_golden.py::disc_case): logits of all six discriminators,
feature-map statistics, the three adversarial losses, parameter gradients of the discriminator step, dL/dy_hat of the generator step, kl_loss."""
import os

import numpy as np
import pytest
import torch

from oracle import disc_oracle as DO


@pytest.fixture(scope="module")
def z(golden_dir):
    return np.load(os.path.join(golden_dir, "disc.npz"))


def test_discriminator_step_matches_reference(z):
    P = {k: v.clone().requires_grad_(True) for k, v in DO.init_params(seed=4).items()}
    y, y_hat = torch.tensor(z["y"]), torch.tensor(z["y_hat"])
    y_d_r, y_d_g, _, _ = DO.mpd(P, y, y_hat)
    for i in range(6):
        for got, key in ((y_d_r[i], "d%d_real" % i), (y_d_g[i], "d%d_gen" % i)):
            want = z[key]
            assert tuple(got.shape) == want.shape
            assert np.abs(got.detach().numpy() - want).max() <= 2e-5 * max(1.0, np.abs(want).max()), key
    loss_d = DO.discriminator_loss(y_d_r, y_d_g)
    assert abs(float(loss_d.detach()) - float(z["loss_d"])) <= 1e-5 * float(z["loss_d"])
    loss_d.backward()
    names = [str(n) for n in z["names"]]
    assert set(names) == set(P.keys())
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = P[k].grad
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 1e-2 * scale + floor, k


def test_generator_side_losses_and_waveform_gradient_match_reference(z):
    P = DO.init_params(seed=4)
    y = torch.tensor(z["y"])
    y_hat = torch.tensor(z["y_hat"]).requires_grad_(True)
    _, y_d_g, fmap_r, fmap_g = DO.mpd(P, y, y_hat)
    loss_fm, loss_gen = DO.feature_loss(fmap_r, fmap_g), DO.generator_loss(y_d_g)
    assert abs(float(loss_fm) - float(z["loss_fm"])) <= 1e-5 * float(z["loss_fm"])
    assert abs(float(loss_gen) - float(z["loss_gen"])) <= 1e-5 * float(z["loss_gen"])
    for d, fm in enumerate(fmap_g):
        assert len(fm) == (7 if d == 0 else 6)                        # scale discriminator: 6 convs + post; period ones: 5 + post
        for j, f in enumerate(fm):
            assert abs(float(f.mean()) - z["fmap_mean"][d, j]) <= 1e-5 + 1e-4 * abs(z["fmap_mean"][d, j])
            assert abs(float(f.abs().mean()) - z["fmap_abs"][d, j]) <= 1e-4 * z["fmap_abs"][d, j]
    (loss_gen + loss_fm).backward()
    want = z["dy_hat"]
    assert np.linalg.norm(y_hat.grad.numpy() - want) <= 1e-4 * np.linalg.norm(want)


def test_period_fold_and_kl_loss(z):
    P = DO.init_params(seed=4)
    # a length that IS a multiple of the period is folded without padding; the logits length follows the three stride-3 convolutions
    x = torch.tensor(z["y"])[:, :, :2 * 3 * 5 * 7 * 11]
    with torch.no_grad():
        for d, p in enumerate(DO.PERIODS):
            out, fm = DO.disc_p(P, x, p, "discriminators.%d." % (d + 1))
            rows = x.shape[-1] // p
            for _ in range(4):
                rows = (rows + 2 * 2 - 5) // 3 + 1
            assert out.shape == (2, rows * p) and fm[0].shape[-1] == p
    a, b, c, dd = [torch.tensor(t) for t in z["kl_in"]]
    assert abs(float(DO.kl_loss(a, b, c, dd, torch.tensor(z["kl_mask"]))) - float(z["kl"])) <= 1e-5 * abs(float(z["kl"]))
