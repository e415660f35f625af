"""CPU emulation of the adversarial-loss reductions (tests/emu builds ttts_b200/csrc/gan_losses.cu for the host) against the op contract
tests/ref_kernels.py (losses.py:7-44 of the reference)."""
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


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libgan_losses_emu.so")
    compile_emu("gan_losses_emu.cpp", so)
    lib = ctypes.CDLL(so)
    vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
    lib.ttts_lsgan_loss.argtypes = [vp, f32, i64, vp, vp, vp]
    lib.ttts_lsgan_loss_bwd.argtypes = [vp, f32, vp, i64, vp, vp]
    lib.ttts_l1_mean.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.ttts_l1_mean_bwd.argtypes = [vp, vp, vp, i64, vp, vp]
    i32 = ctypes.c_int32
    lib.ttts_kl_loss.argtypes = [vp] * 5 + [i32, i32, i32, vp, vp, vp]
    lib.ttts_kl_loss_bwd.argtypes = [vp] * 6 + [i32, i32, i32, vp, vp, vp, vp, vp]
    return lib


@pytest.mark.parametrize("n", [255, 70001])
def test_loss_reductions_and_their_gradients(emu, n):
    g = torch.Generator().manual_seed(n)
    x, a = torch.randn(n, generator=g), torch.randn(n, generator=g)
    scratch, out, dL = torch.zeros(256), torch.zeros(1), torch.tensor([0.7])
    for c in (0.0, 1.0):
        assert emu.ttts_lsgan_loss(x.data_ptr(), c, n, scratch.data_ptr(), out.data_ptr(), None) == 0
        assert abs(float(out) - float(R.lsgan_fwd(x, c))) <= 2e-6 * max(1.0, float(R.lsgan_fwd(x, c)))
        dx = torch.empty(n)
        assert emu.ttts_lsgan_loss_bwd(x.data_ptr(), c, dL.data_ptr(), n, dx.data_ptr(), None) == 0
        assert float((dx - R.lsgan_bwd(dL, x, c)).abs().max()) <= 1e-6
    assert emu.ttts_l1_mean(a.data_ptr(), x.data_ptr(), n, scratch.data_ptr(), out.data_ptr(), None) == 0
    assert abs(float(out) - float(R.l1_fwd(a, x))) <= 2e-6 * max(1.0, float(R.l1_fwd(a, x)))
    db = torch.empty(n)
    assert emu.ttts_l1_mean_bwd(a.data_ptr(), x.data_ptr(), dL.data_ptr(), n, db.data_ptr(), None) == 0
    assert float((db - R.l1_bwd(dL, a, x)).abs().max()) <= 1e-7


def test_kl_loss_and_its_four_gradients(emu):
    g = torch.Generator().manual_seed(9)
    B, C, T = 3, 10, 17
    z_p, logs_q, m_p, logs_p = [0.5 * torch.randn(B, C, T, generator=g) for _ in range(4)]
    mask = (torch.arange(T)[None, :] < torch.tensor([17, 9, 1])[:, None]).float().contiguous()
    scratch, out2, dL = torch.zeros(512), torch.zeros(2), torch.tensor([1.3])
    P = lambda t: t.data_ptr()
    assert emu.ttts_kl_loss(P(z_p), P(logs_q), P(m_p), P(logs_p), P(mask), B, C, T, P(scratch), P(out2), None) == 0
    want = R.kl_fwd(z_p, logs_q, m_p, logs_p, mask)
    assert abs(float(out2[0]) - float(want)) <= 2e-6 * max(1.0, abs(float(want))) and float(out2[1]) == 27.0
    gs = [torch.empty(B, C, T) for _ in range(4)]
    assert emu.ttts_kl_loss_bwd(P(z_p), P(m_p), P(logs_p), P(mask), P(dL), P(out2), B, C, T, P(gs[0]), P(gs[1]), P(gs[2]), P(gs[3]), None) == 0
    for got, ref in zip(gs, R.kl_bwd(dL, z_p, logs_q, m_p, logs_p, mask)):
        assert float((got - ref).abs().max()) <= 1e-6 + 1e-5 * float(ref.abs().max())
    assert emu.ttts_kl_loss_bwd(P(z_p), P(m_p), P(logs_p), P(mask), P(dL), P(out2), B, C, T, P(gs[0]), None, None, None, None) == 0   # optional outputs
