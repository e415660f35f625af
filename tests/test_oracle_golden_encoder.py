"""CPU: the encoder-stack oracle (oracle/encoder_oracle.py) against golden vectors minted from the REAL reference modules
(MelStyleEncoder, PosteriorAudioEncoder, proj, quantizer -- tests/golden/make_golden.py::encoder_case)."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as EO
from oracle import vq_mel_oracle as V


@pytest.fixture(scope="module")
def enc(golden_dir):
    return np.load(os.path.join(golden_dir, "encoder.npz"))


def test_filter_matches_reference(enc):
    np.testing.assert_allclose(EO.kaiser_sinc_filter12().numpy().reshape(-1), enc["filt"], rtol=0, atol=1e-7)


def test_encoder_oracle_matches_reference(enc):
    P = EO.init_params(seed=5)
    wav = torch.tensor(enc["wav"])
    spec = torch.tensor(V.spectrogram(enc["wav"]))
    with torch.no_grad():
        o = EO.encode(P, spec, wav, lengths=torch.tensor(enc["lengths"]), eps=torch.tensor(enc["eps"]), codebook=enc["E"])
    def rel(a, b):
        return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-20)
    assert rel(o["ge"].numpy(), enc["ge"]) < 1e-4
    assert rel(o["m"].numpy(), enc["m"]) < 2e-4
    assert rel(o["logs"].numpy(), enc["logs"]) < 2e-4
    assert rel(o["z"].numpy(), enc["z"]) < 2e-4
    assert rel(o["x"].numpy(), enc["x"]) < 2e-4
    # codes: equal except where the fp64 top-2 margin is tiny (the encoder output itself carries ~1e-4 relative noise)
    xn = np.ascontiguousarray(enc["x"].transpose(0, 2, 1)).reshape(-1, 192)
    margin = V.vq_margin(xn, enc["E"], enc["codes"].reshape(-1))
    flips = o["codes"].reshape(-1) != enc["codes"].reshape(-1)
    assert not np.any(flips & (margin > 1e-3))
    assert flips.sum() <= 2


def test_masked_frames_are_zero(enc):
    z = enc["z"]
    assert np.all(z[1, :, 30:] == 0) and np.all(z[2, :, 17:] == 0)
