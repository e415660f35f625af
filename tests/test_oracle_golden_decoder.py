"""CPU: the decoder oracle (oracle/decoder_oracle.py: HiFi-GAN-style Generator of SynthesizerTrn, the next scope row) against golden vectors
minted from the REAL reference module (tests/golden/make_golden.py::decoder_case): waveform and gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_oracle as DO


@pytest.fixture(scope="module")
def dec(golden_dir):
    return np.load(os.path.join(golden_dir, "decoder.npz"))


def test_decoder_oracle_forward_and_gradients_match_reference(dec):
    P = {k: v.clone().requires_grad_(True) for k, v in DO.init_params(seed=9).items()}
    y = DO.generator(P, torch.tensor(dec["z"]), torch.tensor(dec["g"]))
    assert y.shape == dec["y"].shape == (2, 1, 6 * 640)
    assert np.linalg.norm(y.detach().numpy() - dec["y"]) <= 2e-5 * np.linalg.norm(dec["y"])
    R = torch.randn(y.shape, generator=torch.Generator().manual_seed(32))
    loss = (y * R).sum()
    assert abs(float(loss) - float(dec["loss"])) <= 1e-4 * max(1.0, abs(float(dec["loss"])))
    loss.backward()
    names = [str(n) for n in dec["names"]]
    assert set(names) == set(P.keys())
    floor = 1e-6 * float(np.sqrt((dec["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = P[k].grad
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(dec["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(dec["proj"][i])) <= 1e-2 * scale + floor, k


def test_decoder_upsamples_by_640_and_is_bounded(dec):
    P = DO.init_params(seed=9)
    with torch.no_grad():
        y = DO.generator(P, torch.tensor(dec["z"])[:, :, :3], None)
    assert y.shape == (2, 1, 3 * 640) and float(y.abs().max()) <= 1.0
