"""CPU: host logic of the batched VQ extractor (conditioning, equal-length batching, on-disk format) with a stand-in encoder."""
import os

import torch

from ttts_b200.prepare import extract_vq as X


class FakeEncoder:
    """codes[b, n] = a deterministic function of the clip's own samples only (so batching must not change them)."""
    def __call__(self, wav):
        B, L = wav.shape
        n = L // 1280
        c = (wav.reshape(B, n, 1280).abs().sum(-1) * 1000).long() % 1024
        return {"codes": c.unsqueeze(0)}


def test_condition_wav_matches_reference_rules():
    w = torch.linspace(-2, 2, 640 * 37 + 5).unsqueeze(0).repeat(2, 1)          # stereo, 37 hops + 5 samples
    c = X.condition_wav(w)
    assert c.shape == (640 * 36,) and c.min() >= -1 and c.max() <= 1            # first channel, even number of hops, clamped
    assert X.condition_wav(torch.zeros(640 * 16 - 1)) is None                   # shorter than 16 hops: skipped


def test_plan_batches_groups_equal_lengths():
    lengths = [2560, None, 1280 * 10, 2560, 2560, 1280 * 10, 1280 * 3]
    b = X.plan_batches(lengths, batch_size=2)
    assert b == [[2, 5], [6], [0, 3], [4]]
    assert sorted(i for bb in b for i in bb) == [0, 2, 3, 4, 5, 6]


def test_extract_writes_reference_format(tmp_path):
    torch.manual_seed(0)
    clips = {"a/x": torch.randn(1, 640 * 40) * 0.3, "a/y": torch.randn(640 * 41 + 17) * 0.3, "b/z": torch.randn(2, 640 * 52) * 2, "b/short": torch.randn(100)}
    paths = [str(tmp_path / k) for k in clips]
    errors = []

    def load(p):
        k = os.path.relpath(p, tmp_path)
        if k == "b/missing":
            raise FileNotFoundError(k)
        return clips[k]
    done = X.extract_vq(paths + [str(tmp_path / "b/missing")], FakeEncoder(), load_fn=load, batch_size=2, device="cpu", on_error=lambda p, e: errors.append(p))
    assert len(errors) == 1 and set(done) == set(paths[:3])
    enc = FakeEncoder()
    for p in paths[:3]:
        got = torch.load(p + ".vq.pth")
        assert isinstance(got, list) and all(isinstance(c, int) for c in got)                       # list[int], as the GPT dataset expects
        w = X.condition_wav(clips[os.path.relpath(p, tmp_path)])
        assert got == enc(w.unsqueeze(0))["codes"][0, 0].tolist() and len(got) == w.shape[0] // 1280
    assert not os.path.exists(paths[3] + ".vq.pth")
