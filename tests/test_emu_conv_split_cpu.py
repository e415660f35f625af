"""CPU emulation of the split-reduction convolution kernel: tests/emu builds ttts_b200/csrc/conv1d_split.cu -- the source nvcc compiles --
for the host and this test compares it with torch's fp32 conv1d + the fused elementwise work of the reference blocks (WN gate, GLU, Mish,
leaky ReLU, residual, scale, mask; ttts/vqvae/modules.py:136-318, 560-566).  What it shows without a GPU: the chunk ranges of the groups,
the (ci, k) bookkeeping started mid-reduction, the lockstep barriers, the partial-tile reduction and the aliasing of the rings.  The
hardware parity test is tests/test_gpu_encoder.py::test_conv1d_split_vs_torch."""
import ctypes
import os
import shutil
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu_build import compile_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libconv_split_emu.so")
    # TTTS_EMU_CXXFLAGS="-g -fsanitize=address" (with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0) turns every
    # out-of-bounds shared / global access of the emulated kernels into a hard error
    compile_emu("conv_split_emu.cpp", so)
    lib = ctypes.CDLL(so)
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.emu_conv1d_split.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, f32, i32, vp, i32, vp, i32, i32]
    return lib


def reference(x, w, bias, stride, dil, pad, pre_lrelu, resid, out_scale, mask, post, cond):
    xin = F.leaky_relu(x, 0.1) if pre_lrelu else x
    y = F.conv1d(xin, w, bias, stride=stride, dilation=dil, padding=pad)
    if post == 1:
        a, g = y.chunk(2, 1)
        y = a * torch.sigmoid(g)
    elif post == 2:
        y = F.mish(y)
    elif post == 3:
        a, g = y.chunk(2, 1)
        if cond is not None:
            ca, cg = cond.chunk(2, 1)
            a, g = a + ca[:, :, None], g + cg[:, :, None]
        y = torch.tanh(a) * torch.sigmoid(g)
    if resid is not None:
        y = y + resid
    y = y * out_scale
    if mask is not None:
        y = y * mask[:, None, :]
    return y


CASES = [
    # B, Cin, T, Cout, K, stride, dil, pad, lrelu, resid, scale, mask, post, cond, groups
    (3, 48, 36, 96, 5, 1, 1, 2, False, False, 1.0, True, 3, True, 4),      # a WN in_layer in miniature: gated, cond, mask; R = 240 = 15 chunks
    (2, 40, 36, 40, 1, 1, 1, 0, False, True, 1.0, True, 0, False, 2),      # WN res_skip (1x1) with residual + mask; R = 40 -> 3 chunks, G = 2
    (2, 24, 50, 40, 11, 1, 3, 15, True, True, 0.5, False, 0, False, 4),    # ResBlock conv: leaky ReLU, dilation, ragged tile (100 positions)
    (2, 20, 61, 24, 7, 2, 1, 3, True, False, 1.0, False, 2, False, 4),     # strided + Mish; R = 140 -> 9 chunks: groups of 3, 3, 3, 0
    (1, 3, 70, 16, 3, 1, 1, 1, False, False, 1.0, False, 1, False, 4),     # GLU, R = 9 -> ONE chunk: three idle groups
]


@pytest.mark.parametrize("case", CASES)
def test_split_conv_on_the_cpu_emulation(emu, case):
    B, Cin, T, Cout, K, stride, dil, pad, lrelu, use_res, scale, use_mask, post, use_cond, groups = case
    g = torch.Generator().manual_seed(sum(case[:8]))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    bias = torch.randn(Cout, generator=g)
    Tout = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
    Ceff = Cout // 2 if post in (1, 3) else Cout
    resid = torch.randn(B, Ceff, Tout, generator=g) if use_res else None
    mask = (torch.rand(B, Tout, generator=g) > 0.3).float() if use_mask else None
    cond = torch.randn(B, Cout, generator=g) if use_cond else None
    want = reference(x, w, bias, stride, dil, pad, lrelu, resid, scale, mask, post, cond)
    y = torch.full((B, Ceff, Tout), 123.0)
    p = lambda t: t.data_ptr() if t is not None else None
    rc = emu.emu_conv1d_split(p(x), p(w), p(bias), p(y), B, Cin, T, Cout, K, stride, dil, pad, int(lrelu), p(resid), scale, 0, p(mask), post,
                              p(cond), Cout if use_cond else 0, groups)
    assert rc == 0
    assert float((y - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))
    # accumulate: y += ...
    y2 = torch.ones(B, Ceff, Tout)
    rc = emu.emu_conv1d_split(p(x), p(w), p(bias), p(y2), B, Cin, T, Cout, K, stride, dil, pad, int(lrelu), p(resid), scale, 1, p(mask), post,
                              p(cond), Cout if use_cond else 0, groups)
    assert rc == 0 and float((y2 - 1.0 - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))
