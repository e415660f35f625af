"""CPU emulation of the prior encoder's two kernels (tests/emu builds ttts_b200/csrc/text_encoder_kernels.cu for the host) against the op
contract tests/ref_kernels.py: attention with / without the relative-position window, self and cross, ragged lengths; channel LayerNorm."""
import ctypes
import os
import shutil
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu_build import compile_emu  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402

R = TorchRefKernels()
vp, i32 = ctypes.c_void_p, ctypes.c_int32


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libte_emu.so")
    compile_emu("text_encoder_emu.cpp", so)
    lib = ctypes.CDLL(so)
    lib.ttts_attn_small.argtypes = [vp] * 8 + [i32] * 6 + [vp]
    lib.ttts_attn_small_bwd.argtypes = [vp] * 13 + [i32] * 6 + [vp]
    lib.ttts_layernorm_c.argtypes = [vp] * 5 + [i32] * 3 + [vp]
    lib.ttts_layernorm_c_bwd.argtypes = [vp] * 8 + [i32] * 3 + [vp]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def P(t):
    return t.data_ptr() if t is not None else None


def close(got, want, tol=3e-5):
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= tol * max(1.0, float(want.abs().max())), float((got - want).abs().max())


@pytest.mark.parametrize("B,heads,dk,Tq,Tk,win,qlen,klen", [
    (3, 2, 16, 24, 24, 4, [24, 17, 6], [24, 17, 6]),       # encoder self-attention with the relative window, ragged
    (2, 4, 8, 10, 19, 0, [10, 4], [19, 3]),                # MRTE cross attention: queries = frames, keys = text
    (1, 1, 12, 5, 5, 4, [5], [5]),                         # window wider than the sequence
])
def test_attention(emu, B, heads, dk, Tq, Tk, win, qlen, klen):
    g = torch.Generator().manual_seed(B * 100 + Tk)
    C = heads * dk
    q, k, v, do = torch.randn(B, C, Tq, generator=g), torch.randn(B, C, Tk, generator=g), torch.randn(B, C, Tk, generator=g), torch.randn(B, C, Tq, generator=g)
    ek = ev = None
    if win:
        ek, ev = torch.randn(1, 2 * win + 1, dk, generator=g) * dk ** -0.5, torch.randn(1, 2 * win + 1, dk, generator=g) * dk ** -0.5
    ql, kl = torch.tensor(qlen, dtype=torch.int64), torch.tensor(klen, dtype=torch.int64)
    out = torch.empty(B, C, Tq)
    assert emu.ttts_attn_small(P(q), P(k), P(v), P(ek), P(ev), P(ql), P(kl), P(out), B, C, Tq, Tk, heads, win, None) == 0, emu.emu_last_error()
    close(out, R.attn_fwd(q, k, v, ek, ev, ql, kl, heads))
    dq, dk_, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dek = torch.zeros_like(ek) if win else None
    dev = torch.zeros_like(ev) if win else None
    rc = emu.ttts_attn_small_bwd(P(do), P(q), P(k), P(v), P(ek), P(ev), P(ql), P(kl), P(dq), P(dk_), P(dv), P(dek), P(dev), B, C, Tq, Tk, heads, win, None)
    assert rc == 0, emu.emu_last_error()
    wq, wk, wv, wek, wev = R.attn_bwd(do, q, k, v, ek, ev, ql, kl, heads)
    close(dq, wq); close(dk_, wk); close(dv, wv)
    if win:
        close(dek, wek); close(dev, wev)


@pytest.mark.parametrize("B,C,T", [(3, 192, 24), (2, 20, 70), (1, 5, 1)])
def test_channel_layer_norm(emu, B, C, T):
    g = torch.Generator().manual_seed(C)
    x, dy = 2 * torch.randn(B, C, T, generator=g) + 0.5, torch.randn(B, C, T, generator=g)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    y, stats = torch.empty_like(x), torch.empty(B * T * 2)
    assert emu.ttts_layernorm_c(P(x), P(gamma), P(beta), P(y), P(stats), B, C, T, None) == 0
    close(y, R.lnc_fwd(x, gamma, beta))
    dx, dg, db, scratch = torch.empty_like(x), torch.empty(C), torch.empty(C), torch.empty(B * T * 2)
    assert emu.ttts_layernorm_c_bwd(P(dy), P(x), P(stats), P(gamma), P(dx), P(dg), P(db), P(scratch), B, C, T, None) == 0
    wx, wg, wb = R.lnc_bwd(dy, x, gamma, beta)
    close(dx, wx, 1e-4); close(dg, wg, 1e-4); close(db, wb, 1e-4)
