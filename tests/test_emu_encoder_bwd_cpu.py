"""CPU emulation of the training kernels of the encode half (tests/emu builds ttts_b200/csrc/encoder_bwd.cu for the host; its extern "C"
entry points are called on host pointers) against the per-op contract tests/ref_kernels.py -- the same contract the training graph is pinned
with against the REAL reference's gradients (tests/test_train_encoder_cpu.py)."""
import ctypes
import math
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
from ttts_b200.vqvae.train_encoder import kaiser_sinc_filter12  # noqa: E402

R = TorchRefKernels()
vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libenc_bwd_emu.so")
    compile_emu("encoder_bwd_emu.cpp", so)
    lib = ctypes.CDLL(so)
    lib.ttts_ew_add.argtypes = [vp, vp, vp, i64, vp]
    lib.ttts_ew_scale.argtypes = [vp, f32, vp, i64, vp]
    lib.ttts_ew_mul_mask.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.ttts_glu.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.ttts_mish.argtypes = [vp, vp, vp, i64, i32, vp]
    lib.ttts_wn_gate.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.ttts_weight_norm_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    lib.ttts_lrelu.argtypes = [vp, vp, vp, i64, f32, i32, vp]
    lib.ttts_tanh.argtypes = [vp, vp, vp, i64, i32, vp]
    lib.ttts_add_bcast.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.ttts_sum_t.argtypes = [vp, vp, i32, i32, vp]
    lib.ttts_snake_aa_bwd.argtypes = [vp] * 8 + [i32, i32, i32, vp]
    lib.ttts_mha_small_bwd.argtypes = [vp] * 8 + [i32, i32, i32, i32, f32, vp]
    lib.ttts_masked_mean_bwd.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.ttts_posterior_sample_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def P(t):
    return t.data_ptr() if t is not None else None


def close(got, want, tol=2e-5):
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= tol * max(1.0, float(want.abs().max())), float((got - want).abs().max())


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed + sum(shape)))


def test_elementwise(emu):
    a, b = rnd(3, 5, 37), rnd(3, 5, 37, seed=1)
    o = torch.empty_like(a)
    assert emu.ttts_ew_add(P(a), P(b), P(o), a.numel(), None) == 0
    close(o, R.add(a, b), 0)
    assert emu.ttts_ew_scale(P(a), 1.0 / 3.0, P(o), a.numel(), None) == 0
    close(o, R.scale(a, 1.0 / 3.0), 1e-7)
    mask = (torch.rand(3, 37, generator=torch.Generator().manual_seed(2)) > 0.4).float()
    assert emu.ttts_ew_mul_mask(P(a), P(mask), P(o), 3, 5, 37, None) == 0
    close(o, R.mul_mask(a, mask), 0)


def test_glu_mish_gate(emu):
    raw, dy = rnd(2, 12, 19), rnd(2, 6, 19, seed=3)
    y = torch.empty(2, 6, 19)
    assert emu.ttts_glu(P(raw), None, P(y), 2, 6, 19, 0, None) == 0
    close(y, R.glu_fwd(raw))
    draw = torch.empty_like(raw)
    assert emu.ttts_glu(P(raw), P(dy), P(draw), 2, 6, 19, 1, None) == 0
    close(draw, R.glu_bwd(dy, raw))
    x = 3 * rnd(4, 7, 11, seed=5)
    x[0, 0, 0] = 25.0                                                   # beyond the softplus switch-over
    o = torch.empty_like(x)
    assert emu.ttts_mish(P(x), None, P(o), x.numel(), 0, None) == 0
    close(o, R.mish_fwd(x))
    d = rnd(4, 7, 11, seed=6)
    assert emu.ttts_mish(P(x), P(d), P(o), x.numel(), 1, None) == 0
    close(o, R.mish_bwd(d, x))
    for cond in (rnd(3, 16, seed=7), None):
        raw, dy = rnd(3, 16, 40, seed=8), rnd(3, 8, 40, seed=9)
        y = torch.empty(3, 8, 40)
        assert emu.ttts_wn_gate(P(raw), P(cond), None, P(y), None, 3, 8, 40, 0, None) == 0
        close(y, R.gate_fwd(raw, cond))
        draw, dcond = torch.empty_like(raw), torch.empty(3, 16)
        assert emu.ttts_wn_gate(P(raw), P(cond), P(dy), P(draw), P(dcond) if cond is not None else None, 3, 8, 40, 1, None) == 0
        wr, wc = R.gate_bwd(dy, raw, cond)
        close(draw, wr)
        if cond is not None:
            close(dcond, wc)


def test_lrelu_tanh_broadcast_add_and_row_sum(emu):
    x, dy = rnd(3, 4, 21), rnd(3, 4, 21, seed=1)
    o = torch.empty_like(x)
    for slope in (0.1, 0.01):
        assert emu.ttts_lrelu(P(x), None, P(o), x.numel(), slope, 0, None) == 0
        close(o, R.lrelu_fwd(x, slope), 1e-7)
        assert emu.ttts_lrelu(P(x), P(dy), P(o), x.numel(), slope, 1, None) == 0
        close(o, R.lrelu_bwd(dy, x, slope), 1e-7)
    assert emu.ttts_tanh(P(x), None, P(o), x.numel(), 0, None) == 0
    close(o, R.tanh_fwd(x))
    assert emu.ttts_tanh(P(x), P(dy), P(o), x.numel(), 1, None) == 0
    close(o, R.tanh_bwd(dy, x))
    c, bias = rnd(3, 4, 1, seed=2), rnd(4, seed=3)
    assert emu.ttts_add_bcast(P(x), P(c), P(o), 3, 4, 21, 1, None) == 0
    close(o, R.add_bcast_fwd(x, c), 0)
    assert emu.ttts_add_bcast(P(x), P(bias), P(o), 3, 4, 21, 0, None) == 0
    close(o, x + bias.view(1, -1, 1), 0)
    s = torch.empty(3, 4, 1)
    assert emu.ttts_sum_t(P(dy), P(s), 12, 21, None) == 0
    close(s, R.add_bcast_bwd(dy))


def test_weight_norm_backward(emu):
    for shape in ((7, 5, 3), (130, 192, 5), (4, 1, 1)):
        v, dw = rnd(*shape), rnd(*shape, seed=1)
        g = torch.rand(shape[0], 1, 1, generator=torch.Generator().manual_seed(2)) + 0.5
        dv, dg = torch.empty_like(v), torch.empty(shape[0])
        assert emu.ttts_weight_norm_bwd(P(dw), P(v), P(g), P(dv), P(dg), shape[0], shape[1] * shape[2], None) == 0
        wv, wg = R.wn_bwd(dw, v, g.view(-1))
        close(dv, wv)
        close(dg, wg.view(-1))


@pytest.mark.parametrize("B,C,T", [(2, 5, 36), (1, 3, 7), (2, 2, 70)])
def test_snake_backward(emu, B, C, T):
    x, dy = 1.5 * rnd(B, C, T), rnd(B, C, T, seed=1)
    la, lb = 0.3 * rnd(C, seed=2), 0.3 * rnd(C, seed=3)
    filt = kaiser_sinc_filter12("cpu")
    dx, dla, dlb = torch.empty_like(x), torch.zeros(C), torch.zeros(C)
    assert emu.ttts_snake_aa_bwd(P(dy), P(x), P(la), P(lb), P(filt), P(dx), P(dla), P(dlb), B, C, T, None) == 0, emu.emu_last_error()
    wx, wa, wb = R.snake_bwd(dy, x, la, lb, filt)
    close(dx, wx, 5e-5)
    close(dla, wa, 5e-5)
    close(dlb, wb, 5e-5)


@pytest.mark.parametrize("B,heads,dk,T,lens", [(3, 2, 64, 36, [36, 30, 17]), (2, 2, 8, 5, [5, 1]), (1, 1, 16, 64, [50])])
def test_mha_backward(emu, B, heads, dk, T, lens):
    C = heads * dk
    q, k, v, do = [rnd(B, C, T, seed=s) * (0.5 if s < 2 else 1.0) for s in range(4)]
    lens_t = torch.tensor(lens, dtype=torch.int64)
    temp = math.sqrt(128.0)
    dq, dk_, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    assert emu.ttts_mha_small_bwd(P(do), P(q), P(k), P(v), P(lens_t), P(dq), P(dk_), P(dv), B, C, T, heads, temp, None) == 0, emu.emu_last_error()
    wq, wk, wv = R.mha_bwd(do, q, k, v, lens_t, heads, temp)
    close(dq, wq)
    close(dk_, wk)
    close(dv, wv)


def test_masked_mean_and_posterior_backward(emu):
    lens = torch.tensor([9, 4, 1], dtype=torch.int64)
    dy = rnd(3, 6)
    dx = torch.empty(3, 6, 9)
    assert emu.ttts_masked_mean_bwd(P(dy), P(lens), P(dx), 3, 6, 9, None) == 0
    close(dx, R.masked_mean_bwd(dy, lens, 9), 1e-7)
    stats, eps, dz = rnd(3, 8, 9), rnd(3, 4, 9, seed=1), rnd(3, 4, 9, seed=2)
    mask = (torch.arange(9)[None, :] < lens[:, None]).float()
    ds = torch.empty_like(stats)
    assert emu.ttts_posterior_sample_bwd(P(dz), P(stats), P(eps), P(mask), P(ds), 3, 4, 9, None) == 0
    close(ds, R.posterior_bwd(dz, stats, eps, mask))
    assert emu.ttts_posterior_sample_bwd(P(dz), P(stats), None, P(mask), P(ds), 3, 4, 9, None) == 0
    close(ds, R.posterior_bwd(dz, stats, None, mask))
