"""The diffusion oracle (oracle/diffusion_oracle.py) against the REAL reference's AA_diffusion + SpacedDiffusion.training_losses micro-step
(tests/golden/diffusion.npz, minted by tests/golden/make_golden.py::diffusion_case in train() mode with the random decisions pinned)."""
import os

import numpy as np
import torch

from oracle import diffusion_oracle as DO


def _run():
    cfg = DO.default_config(**DO.GOLDEN_CFG)
    P = {k: v.clone().requires_grad_(True) for k, v in DO.init_params(cfg, seed=12).items()}
    I = DO.golden_inputs()
    x_t = DO.q_sample(I["x_start"], I["t"], I["noise"])
    out = DO.model_forward(P, cfg, x_t, torch.tensor(I["t"]), I["latent"], I["refer"], I["uncond"], I["dropped"])
    mse, vb = DO.loss_terms(out, I["x_start"], x_t, I["noise"], I["t"])
    return P, out, mse, vb


def test_diffusion_oracle_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "diffusion.npz"))
    P, out, mse, vb = _run()
    assert np.abs(out.detach().numpy() - z["model_out"]).max() <= 2e-5 * np.abs(z["model_out"]).max()
    assert np.allclose(mse.detach().numpy(), z["mse"], rtol=2e-6, atol=0)
    assert np.allclose(vb.detach().numpy(), z["vb"], rtol=2e-5, atol=1e-9)
    loss = (mse + vb).mean()
    assert abs(float(loss) - float(z["loss"])) <= 2e-6 * abs(float(z["loss"]))
    loss.backward()
    names = [str(n) for n in z["names"]]
    assert set(names) == set(P.keys())
    floor = 1e-7 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-4 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 2e-4 * scale + floor, k


def test_bucket_table_and_nearest_index():
    b = DO.rel_pos_bucket(200, 200)
    assert int(b.min()) == 0 and int(b.max()) == 31
    # a function of (j - i) only: the kernels index it by the diagonal
    assert torch.equal(b[5:, 5:], b[:-5, :-5])
    for tin, tout in ((6, 24), (7, 24), (256, 1024), (100, 400), (33, 100)):
        x = torch.arange(tin, dtype=torch.float32)[None, None]
        ref = torch.nn.functional.interpolate(x, size=tout, mode="nearest")[0, 0].long()
        assert torch.equal(ref, DO.nearest_index(tin, tout)), (tin, tout)
