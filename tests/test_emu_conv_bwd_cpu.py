"""CPU emulation of the Conv1d backward kernels (tests/emu builds ttts_b200/csrc/conv1d_bwd.cu for the host): input / weight / bias gradients
against torch.autograd of `F.conv1d(F.leaky_relu(x, 0.1), w, b, stride, padding, dilation)` -- the op ttts_conv1d_f32 computes forward
(nn.Conv1d of the reference's encoder, ttts/vqvae/vq2.py:667-745, modules.py:136-318).  First kernels of the next scope row; the hardware
parity test is tests/test_gpu_encoder.py::test_conv1d_backward_vs_autograd."""
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
    so = str(tmp_path_factory.mktemp("emu") / "libconv_bwd_emu.so")
    compile_emu("conv_bwd_emu.cpp", so)
    lib = ctypes.CDLL(so)
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    lib.emu_conv1d_bwd_input.argtypes = [vp, vp, vp, vp] + [i32] * 10
    lib.emu_conv1d_bwd_weight.argtypes = [vp, vp, vp, vp] + [i32] * 9
    lib.emu_bias_grad.argtypes = [vp, vp, i32, i32, i32]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


CASES = [
    # B, Cin, T, Cout, K, stride, dil, pad, lrelu
    (2, 24, 50, 40, 11, 1, 3, 15, True),       # ResBlock conv (dilated, "same"), ragged tiles
    (3, 48, 36, 96, 5, 1, 1, 2, False),        # WN in_layer in miniature
    (2, 16, 160, 32, 16, 8, 1, 7, False),      # a downsampling conv (kernel 2 x stride): dgrad hits the divisibility test
    (2, 20, 61, 24, 7, 2, 1, 3, True),         # strided + leaky ReLU, odd length
    (1, 3, 70, 5, 3, 1, 1, 1, False),          # channel counts far below the tile
    (2, 40, 36, 40, 1, 1, 1, 0, False),        # 1x1 (WN res_skip, Linear)
    (1, 3, 700, 5, 3, 1, 1, 1, True),          # long sequence: several slices of the position axis (atomic combination), many chunks
]


@pytest.mark.parametrize("case", CASES)
def test_conv1d_backward_on_the_cpu_emulation(emu, case):
    B, Cin, T, Cout, K, stride, dil, pad, lrelu = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, T, generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g, requires_grad=True)
    y = F.conv1d(F.leaky_relu(x, 0.1) if lrelu else x, w, b, stride=stride, dilation=dil, padding=pad)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    p = lambda t: t.data_ptr() if t is not None else None
    xd, wd = x.detach().contiguous(), w.detach().contiguous()
    # input gradient (plain, then accumulating on top of a known tensor)
    dx = torch.full_like(xd, 77.0)
    assert emu.emu_conv1d_bwd_input(p(dy), p(wd), p(xd), p(dx), B, Cin, T, Cout, K, stride, dil, pad, int(lrelu), 0) == 0, emu.emu_last_error()
    scale = max(1.0, float(x.grad.abs().max()))
    assert float((dx - x.grad).abs().max()) <= 2e-5 * scale
    dx2 = torch.ones_like(xd)
    assert emu.emu_conv1d_bwd_input(p(dy), p(wd), p(xd) if lrelu else None, p(dx2), B, Cin, T, Cout, K, stride, dil, pad, int(lrelu), 1) == 0
    assert float((dx2 - 1.0 - x.grad).abs().max()) <= 2e-5 * scale
    # weight / bias gradients accumulate into what is there
    dw = torch.full_like(wd, 0.5)
    db = torch.full((Cout,), -2.0)
    assert emu.emu_conv1d_bwd_weight(p(dy), p(xd), p(dw), p(db), B, Cin, T, Cout, K, stride, dil, pad, int(lrelu)) == 0, emu.emu_last_error()
    wscale = max(1.0, float(w.grad.abs().max()))
    assert float((dw - 0.5 - w.grad).abs().max()) <= 5e-5 * wscale
    assert float((db + 2.0 - b.grad).abs().max()) <= 5e-5 * max(1.0, float(b.grad.abs().max()))
    # no bias gradient requested
    dw2 = torch.zeros_like(wd)
    assert emu.emu_conv1d_bwd_weight(p(dy), p(xd), p(dw2), None, B, Cin, T, Cout, K, stride, dil, pad, int(lrelu)) == 0
    assert float((dw2 - w.grad).abs().max()) <= 5e-5 * wscale


@pytest.mark.parametrize("Cin,Cout,K,stride,pad,T", [(12, 6, 16, 10, 3, 7), (8, 4, 16, 8, 4, 9), (6, 3, 8, 2, 3, 20), (5, 40, 2, 2, 0, 33)])
def test_transposed_convolution_from_the_same_kernels(emu, Cin, Cout, K, stride, pad, T):
    """ConvTranspose1d (the Generator's up-sampling layers, vq2.py:369-378: kernel / stride pairs (16,10) (16,8) (8,2) (2,2)) needs no kernel
    of its own: forward = the dgrad kernel, weight gradient = the wgrad kernel with input and output-gradient swapped (dx = the forward
    convolution, GPU-validated already).  Checked against torch's conv_transpose1d and its autograd."""
    B = 2
    g = torch.Generator().manual_seed(Cin * 131 + K)
    x = torch.randn(B, Cin, T, generator=g, requires_grad=True)
    w = (torch.randn(Cin, Cout, K, generator=g) / (Cin * K) ** 0.5).requires_grad_(True)
    y = F.conv_transpose1d(x, w, None, stride=stride, padding=pad)
    Tout = (T - 1) * stride - 2 * pad + K
    assert y.shape[-1] == Tout
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    p = lambda t: t.data_ptr() if t is not None else None
    xd, wd = x.detach().contiguous(), w.detach().contiguous()
    got = torch.full((B, Cout, Tout), 9.0)
    assert emu.emu_conv1d_bwd_input(p(xd), p(wd), None, p(got), B, Cout, Tout, Cin, K, stride, 1, pad, 0, 0) == 0, emu.emu_last_error()
    assert float((got - y.detach()).abs().max()) <= 2e-5 * max(1.0, float(y.detach().abs().max()))
    dw = torch.zeros_like(wd)
    assert emu.emu_conv1d_bwd_weight(p(xd), p(dy), p(dw), None, B, Cout, Tout, Cin, K, stride, 1, pad, 0) == 0, emu.emu_last_error()
    assert float((dw - w.grad).abs().max()) <= 5e-5 * max(1.0, float(w.grad.abs().max()))
    db = torch.full((Cout,), 1.5)
    assert emu.emu_bias_grad(p(dy), p(db), B, Cout, Tout) == 0
    assert float((db - 1.5 - dy.sum(dim=(0, 2))).abs().max()) <= 5e-5 * max(1.0, float(dy.sum(dim=(0, 2)).abs().max()))
    # dx of the transposed convolution is the plain forward convolution of dy with the same weight
    assert float((F.conv1d(dy, wd, None, stride=stride, padding=pad) - x.grad).abs().max()) <= 2e-5 * max(1.0, float(x.grad.abs().max()))


@pytest.mark.parametrize("B,C,T", [(3, 5, 2052), (2, 4, 1024), (5, 3, 8), (2, 3, 1030), (1, 2, 7)])
def test_bias_gradient_kernels(emu, B, C, T):
    """ttts_bias_grad: the 16-byte-load kernel (T a multiple of 4: items of 1024 positions dealt out over the slices, four loads in flight, the
    ragged last item of a row) and the scalar kernel (any T); both ADD to db"""
    g = torch.Generator().manual_seed(B * 100 + T)
    dy = torch.randn(B, C, T, generator=g)
    db = torch.full((C,), -2.25)
    assert emu.emu_bias_grad(dy.data_ptr(), db.data_ptr(), B, C, T) == 0, emu.emu_last_error()
    want = dy.double().sum(dim=(0, 2))
    assert float((db.double() + 2.25 - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))


def test_bad_arguments_are_reported(emu):
    t = torch.zeros(8)
    assert emu.emu_conv1d_bwd_input(t.data_ptr(), t.data_ptr(), None, t.data_ptr(), 1, 1, 2, 1, 5, 1, 1, 0, 0, 0) != 0     # empty output
    assert b"empty output" in emu.emu_last_error()
    assert emu.emu_conv1d_bwd_input(t.data_ptr(), t.data_ptr(), None, t.data_ptr(), 1, 1, 8, 1, 1, 1, 1, 0, 1, 0) != 0     # pre_lrelu without x
