"""CPU: the prior-encoder oracle (oracle/text_encoder_oracle.py: TextEncoder + MRTE with the windowed relative-position attention restated in
closed form; the next scope row) against golden vectors minted from the REAL reference module (make_golden.py::text_encoder_case)."""
import os

import numpy as np
import torch

from oracle import text_encoder_oracle as TO


def test_text_encoder_oracle_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "text_encoder.npz"))
    y, y_lengths, text, text_lengths, ge = TO.golden_inputs()
    assert abs(float(y.sum()) - float(z["y_sum"])) < 1e-3
    P = {k: v.clone().requires_grad_(True) for k, v in TO.init_params(seed=8).items()}
    y.requires_grad_(True); ge.requires_grad_(True)
    yo, m, logs = TO.text_encoder(P, y, y_lengths, text, text_lengths, ge)
    for got, key in ((yo, "yo"), (m, "m"), (logs, "logs")):
        assert np.abs(got.detach().numpy() - z[key]).max() <= 5e-5 * max(1.0, np.abs(z[key]).max()), key
    gR = torch.Generator().manual_seed(62)
    R1, R2 = torch.randn(m.shape, generator=gR), torch.randn(m.shape, generator=gR)
    loss = (m * R1).sum() + (logs * R2).sum()
    assert abs(float(loss.detach()) - float(z["loss"])) <= 1e-4 * max(1.0, abs(float(z["loss"])))
    loss.backward()
    names = [str(n) for n in z["names"]]
    assert set(names) == set(P.keys())
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 1e-2 * scale + floor, k
    assert np.linalg.norm(y.grad.numpy() - z["dy"]) <= 1e-4 * np.linalg.norm(z["dy"])
    assert np.linalg.norm(ge.grad.numpy() - z["dge"]) <= 1e-4 * np.linalg.norm(z["dge"])


def test_masked_positions_do_not_leak():
    """frames / tokens beyond the lengths must not influence the valid outputs (the -1e4 masking + the x_mask products)"""
    y, y_lengths, text, text_lengths, ge = TO.golden_inputs()
    P = TO.init_params(seed=8)
    with torch.no_grad():
        _, m0, _ = TO.text_encoder(P, y, y_lengths, text, text_lengths, ge)
        y2, text2 = y.clone(), text.clone()
        y2[1, :, 17:] += 5.0
        text2[2, 3:] = 7
        _, m1, _ = TO.text_encoder(P, y2, y_lengths, text2, text_lengths, ge)
    assert float((m0 - m1).abs().max()) <= 1e-4
    assert float(m0[2, :, 6:].abs().max()) == 0.0
