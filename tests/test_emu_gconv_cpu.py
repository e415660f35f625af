"""CPU emulation of the grouped Conv1d kernels (tests/emu builds ttts_b200/csrc/conv1d_grouped.cu for the host) against torch and its
autograd, on DiscriminatorS-like layers (ttts/vqvae/vq2.py:498-507: kernel 41, stride 4, four input channels per group)."""
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
    so = str(tmp_path_factory.mktemp("emu") / "libgconv_emu.so")
    compile_emu("gconv_emu.cpp", so)
    lib = ctypes.CDLL(so)
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    lib.ttts_gconv1d.argtypes = [vp] * 4 + [i32] * 8 + [vp]
    lib.ttts_gconv1d_bwd.argtypes = [vp] * 5 + [i32] * 8 + [vp]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


@pytest.mark.parametrize("B,Cin,T,Cout,K,stride,pad,groups", [(2, 8, 123, 16, 41, 4, 20, 2), (1, 16, 90, 16, 41, 4, 20, 4), (2, 6, 37, 9, 5, 1, 2, 3),
                                                              (2, 8, 50, 8, 3, 2, 1, 1)])
def test_grouped_conv_forward_and_backward(emu, B, Cin, T, Cout, K, stride, pad, groups):
    g = torch.Generator().manual_seed(Cin + Cout + K)
    x = torch.randn(B, Cin, T, generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin // groups, K, generator=g) / (Cin // groups * K) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, generator=g)
    y = F.conv1d(x, w, b, stride=stride, padding=pad, groups=groups)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    P = lambda t: t.data_ptr() if t is not None else None
    xd, wd = x.detach().contiguous(), w.detach().contiguous()
    got = torch.empty_like(y)
    assert emu.ttts_gconv1d(P(xd), P(wd), P(b), P(got), B, Cin, T, Cout, K, stride, pad, groups, None) == 0, emu.emu_last_error()
    assert float((got - y.detach()).abs().max()) <= 2e-5 * max(1.0, float(y.detach().abs().max()))
    dx, dw = torch.empty_like(xd), torch.full_like(wd, 0.25)
    assert emu.ttts_gconv1d_bwd(P(dy), P(xd), P(wd), P(dx), P(dw), B, Cin, T, Cout, K, stride, pad, groups, None) == 0, emu.emu_last_error()
    assert float((dx - x.grad).abs().max()) <= 2e-5 * max(1.0, float(x.grad.abs().max()))
    assert float((dw - 0.25 - w.grad).abs().max()) <= 5e-5 * max(1.0, float(w.grad.abs().max()))
    assert emu.ttts_gconv1d_bwd(P(dy), P(xd), P(wd), None, None, B, Cin, T, Cout, K, stride, pad, groups, None) == 0      # both outputs optional
    assert emu.ttts_gconv1d(P(xd), P(wd), None, P(got), B, Cin, T, Cout, K, stride, pad, groups + 5, None) != 0            # indivisible groups
