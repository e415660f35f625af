"""CPU: the data edge in front of the GPT step (ttts_b200/gpt/dataset.py) -- manifest, .vq.pth codes, filters, collation -- against the
rules of ttts/gpt/dataset.py."""
import json

import torch

from ttts_b200.gpt import dataset as D


class CharTok:
    def encode(self, txt):
        return [1 + (ord(c) % 200) for c in txt]


def _make(tmp_path, items):
    lines = []
    for name, text, codes in items:
        p = str(tmp_path / name)
        if codes is not None:
            torch.save([int(c) for c in codes], p + ".vq.pth")
        lines.append({"path": p, "text": text})
    man = tmp_path / "data.jsonl"
    D.write_jsonl(str(man), lines)
    return {"dataset": {"path": str(man)}, "dataloader": {"batch_size": 4, "shuffle": False, "num_workers": 0, "drop_last": False}}


def test_jsonl_roundtrip_keeps_unicode(tmp_path):
    p = tmp_path / "m.jsonl"
    rows = [{"path": "a/b.wav", "text": "你好，世界"}, {"path": "c.wav", "text": "x"}]
    D.write_jsonl(str(p), rows)
    assert D.read_jsonl(str(p)) == rows and "你好" in open(p, encoding="utf-8").read()       # ensure_ascii=False like the reference


def test_dataset_items_and_filters(tmp_path):
    cfg = _make(tmp_path, [("a.wav", "hello", range(10)), ("long_text.wav", "x" * 401, range(5)), ("long_mel.wav", "hi", range(601)),
                           ("missing.wav", "hi", None), ("edge.wav", "y" * 400, range(600))])
    ds = D.GptTtsDataset(cfg, tokenizer=CharTok(), text_fn=None, wav_length_fn=lambda p: 24000)
    assert len(ds) == 5
    text, qmel, n = ds[0]
    assert text.dtype == torch.int64 and text.tolist() == CharTok().encode("hello") and qmel.tolist() == list(range(10)) and n == 24000
    assert ds[1] is None and ds[2] is None and ds[3] is None            # > 400 text ids, > 600 codes, unreadable .vq.pth
    assert ds[4] is not None                                            # exactly at the limits is kept


def test_collater_pads_with_zero_and_drops_none(tmp_path):
    col = D.GptTtsCollater({})
    a = (torch.LongTensor([5, 6, 7]), torch.LongTensor([1, 2]), 1000)
    b = (torch.LongTensor([9]), torch.LongTensor([3, 4, 5, 6]), 2500)
    out = col([a, None, b])
    assert set(out) == {"padded_text", "text_lengths", "padded_qmel", "qmel_lengths", "wav_lens"}
    assert out["padded_text"].tolist() == [[5, 6, 7], [9, 0, 0]] and out["padded_qmel"].tolist() == [[1, 2, 0, 0], [3, 4, 5, 6]]
    assert out["text_lengths"].tolist() == [3, 1] and out["qmel_lengths"].tolist() == [2, 4] and out["wav_lens"].tolist() == [1000, 2500]
    assert all(v.dtype == torch.int64 for v in out.values())
    assert col([None, None]) is None


def test_dataloader_end_to_end_and_sharding(tmp_path):
    cfg = _make(tmp_path, [("c%d.wav" % i, "t" * (i + 1), range(i + 2)) for i in range(6)])
    ds = D.GptTtsDataset(cfg, tokenizer=CharTok(), text_fn=None, wav_length_fn=lambda p: 1024 * 7)
    batches = list(D.build_dataloader(cfg, dataset=ds))
    assert [b["padded_text"].shape for b in batches] == [torch.Size([4, 4]), torch.Size([2, 6])]
    assert batches[1]["padded_qmel"].shape == (2, 7) and batches[0]["wav_lens"].tolist() == [7168] * 4
    seen = []
    for r in range(2):
        for b in D.build_dataloader(cfg, rank=r, world=2, dataset=ds):
            seen += b["text_lengths"].tolist()
    assert sorted(seen) == [1, 2, 3, 4, 5, 6]                          # every sample on exactly one rank


def test_text_normalisation_rules():
    f = D.BpeTextTokenizer.preprocess_text
    assert f("a{b}[c]`d—e") == "a(b)(c)'d-e" and f("@") == "" and f("a@b") == "a@b"


def test_tokenizer_matches_reference_class_when_checkout_is_present():
    """With the reference checkout at hand (this container; not the GPU box) the ids equal those of its own VoiceBpeTokenizer."""
    import importlib.util
    import os
    import pytest
    root = "/root/reference/ttts/gpt"
    vocab = os.path.join(root, "gpt_tts_tokenizer.json")
    if not os.path.exists(vocab):
        pytest.skip("reference checkout not present")
    spec = importlib.util.spec_from_file_location("_ref_voice_tokenizer", os.path.join(root, "voice_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref, ours = mod.VoiceBpeTokenizer(vocab), D.BpeTextTokenizer(vocab)
    for txt in ["ni3 hao3 shi4 jie4", "zhe4 shi4 yi1 ge4 ce4 shi4 [x] {y} `z` — ok", "da4 jia1 hao3 , wo3 shi4 ，。？", "@", ""]:
        a, b = ref.encode(txt), ours.encode(txt)
        assert a == b, txt
        assert ref.decode(torch.tensor(a)) == ours.decode(torch.tensor(b))
