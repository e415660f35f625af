"""CPU: host logic of the batched mel extractor (mono mix-down, equal-length batching, on-disk format) with a stand-in front end."""
import os

import torch

from ttts_b200.prepare import extract_mel as X


def fake_extractor(wav):
    """[B, L] -> [B, 4, 1 + L // 256]: depends on the clip's own samples only, so batching must not change it."""
    B, L = wav.shape
    F = 1 + L // 256
    frames = torch.nn.functional.pad(wav, (0, F * 256 - L)).reshape(B, F, 256)
    return torch.stack([frames.sum(-1), frames.abs().sum(-1), frames.max(-1).values, frames.min(-1).values], dim=1)


def test_condition_wav_mixes_channels_by_mean():
    w = torch.stack([torch.ones(2000), 3 * torch.ones(2000)])
    assert torch.equal(X.condition_wav(w), 2 * torch.ones(2000))                  # mel_extract.py:19-20 (mean), not "first channel"
    assert X.condition_wav(torch.zeros(1, 300)) is None and X.condition_wav(torch.zeros(512)) is None
    assert X.condition_wav(torch.zeros(1, 513)).shape == (513,)


def test_plan_batches_groups_equal_lengths():
    b = X.plan_batches([1000, None, 3000, 1000, 1000, 3000, 700], batch_size=2)
    assert b == [[2, 5], [0, 3], [4], [6]]


def test_extract_writes_reference_format(tmp_path):
    torch.manual_seed(0)
    clips = {"a/x.wav": torch.randn(1, 24000), "a/y.wav": torch.randn(24000), "b/z.wav": torch.randn(2, 31111), "b/short.wav": torch.randn(100)}
    paths = [str(tmp_path / k) for k in clips]
    errors = []

    def load(p):
        k = os.path.relpath(p, tmp_path)
        if k == "b/missing.wav":
            raise FileNotFoundError(k)
        return clips[k]
    done = X.extract_mel(paths + [str(tmp_path / "b/missing.wav")], fake_extractor, load_fn=load, batch_size=2, device="cpu", on_error=lambda p, e: errors.append(p))
    assert len(errors) == 1 and set(done) == set(paths[:3])
    for p in paths[:3]:
        got = torch.load(p + ".mel.pth")
        w = X.condition_wav(clips[os.path.relpath(p, tmp_path)])
        want = fake_extractor(w.unsqueeze(0))
        assert got.shape == want.shape == (1, 4, 1 + w.shape[0] // 256) and got.device.type == "cpu"      # [1, n_mels, frames], as extract_vq.py:13 loads it
        assert torch.equal(got, want) and done[p] == want.shape[-1]
    assert not os.path.exists(paths[3] + ".mel.pth")
