"""CPU: the C-ABI library builds/loads and exports every symbol include/ttts_b200.h declares; host-side module logic
(state_dict parity with the reference, flat-buffer plumbing, loud failure without a GPU)."""
import copy
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    so = g.build()
    return ctypes.CDLL(so)


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ttts_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ttts_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n
    assert lib.ttts_version() >= 100


def test_param_layout_is_consistent():
    from ttts_b200.gpt import engine as E
    from oracle import gpt_oracle as O
    cfg = E.GptConfig(layers=3, model_dim=256, heads=4, max_text_tokens=50, max_mel_tokens=70, n_text_vocab=257, n_mel_vocab=1026,
                      start_text_token=255, stop_text_token=0, start_mel_token=1024, stop_mel_token=1025, mel_length_compression=1024)
    lay = E.Layout(cfg)
    spans = sorted((off, off + n) for _, off, n, _ in lay.entries)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0 and b0 % 64 == 0
    assert spans[-1][1] <= lay.total
    # backward stages tile the flat buffer exactly once
    ranges = sorted(lay.stage_range(s) for s in range(cfg.layers + 2))
    assert ranges[0][0] == 0 and ranges[-1][1] == lay.total
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0
    ocfg = O.default_config(layers=3, model_dim=256, heads=4, max_text_tokens=50, max_mel_tokens=70)
    assert [(n, tuple(s)) for n, _, _, s in lay.entries] == [(n, tuple(s)) for n, s in O.param_shapes(ocfg)]


def test_module_state_dict_and_flat_views():
    from ttts_b200.gpt.model import UnifiedVoice
    from oracle import gpt_oracle as O
    cfg = O.default_config(layers=2, model_dim=128, heads=2, max_text_tokens=40, max_mel_tokens=80)
    kw = {k: cfg[k] for k in ("layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "number_text_tokens", "start_text_token",
                              "number_mel_codes", "start_mel_token", "stop_mel_token")}
    m = UnifiedVoice(**kw)
    sd = m.state_dict()
    assert list(sd.keys()) == [n for n, _ in O.param_shapes(cfg)]
    p = O.init_params(cfg, 0)
    m.load_state_dict(p)
    v = m._layout.views(m._flat)
    assert all(torch.equal(v[k], p[k]) for k in p)
    # optimizers see ordinary parameters; in-place updates land in the flat buffer
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    for q in m.parameters():
        q.grad = torch.ones_like(q)
    before = m._flat.clone()
    opt.step()
    assert not torch.equal(before, m._flat)
    m2 = copy.deepcopy(m)
    assert torch.equal(m2._flat, m._flat) and m2._flat.data_ptr() != m._flat.data_ptr()
    # init statistics follow the reference
    m3 = UnifiedVoice(**kw)
    assert abs(m3.text_embedding.weight.std().item() - 0.02) < 0.003
    assert torch.all(m3.gpt.h[0].ln_1.weight == 1) and torch.all(m3.gpt.h[0].attn.c_attn.bias == 0)


def test_no_cpu_fallback():
    from ttts_b200 import _lib as L
    from ttts_b200.gpt.model import UnifiedVoice
    from ttts_b200.vqvae.quantize import ResidualVectorQuantizer
    from ttts_b200.vqvae.mel import spectrogram_torch
    m = UnifiedVoice(layers=1, model_dim=128, heads=2, max_text_tokens=8, max_mel_tokens=8, number_text_tokens=256, start_text_token=255,
                     number_mel_codes=1026, start_mel_token=1024, stop_mel_token=1025)
    with pytest.raises(L.TTTSError):
        m(torch.zeros(1, 4, dtype=torch.long), torch.tensor([4]), torch.zeros(1, 8, dtype=torch.long), torch.tensor([8192]))
    with pytest.raises(L.TTTSError):
        ResidualVectorQuantizer(dimension=192, n_q=1, bins=1024).encode(torch.zeros(1, 192, 4))
    with pytest.raises(L.TTTSError):
        spectrogram_torch(torch.zeros(1, 4096), 2048, 640, 2048)


def test_sparse_mel_basis_roundtrip():
    import numpy as np
    from ttts_b200.vqvae import mel as M
    from oracle import vq_mel_oracle as V
    b = M.slaney_mel_basis(32000, 2048, 128, 0, None)
    np.testing.assert_allclose(b, V.mel_basis_slaney(), atol=1e-7)
    lo, off, w = M.sparsify(b)
    dense = np.zeros_like(b)
    for m in range(b.shape[0]):
        n = off[m + 1] - off[m]
        dense[m, lo[m]:lo[m] + n] = w[off[m]:off[m + 1]]
    assert np.array_equal(dense, b)
    assert w.size < 0.03 * b.size          # ~98 % sparse
    b24 = M.htk_mel_basis(24000, 1024, 100, 0.0, None)
    np.testing.assert_allclose(b24, V.mel_basis_htk(), atol=1e-7)
