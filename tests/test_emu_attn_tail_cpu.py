"""CPU emulation of the attention tail kernels: tests/emu builds ttts_b200/csrc/attention_tail.cu -- the same source nvcc compiles -- for the
host (one OS thread per CUDA thread, real barriers) and this test runs it against plain torch attention (HF: modeling_gpt2.py:185-226) with the
oracle's restatement of the dropout hash as the keep mask.  It checks what can be checked without a GPU: index arithmetic, the causal range of
every tail row, the dropout keys, the read-add-write of the dK / dV rows.  The parity test of the product path (tile kernels + tail kernels
through the C ABI) is tests/test_gpu_kernels.py::test_attention_* on a B200."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import gpt_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu_build import compile_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libattn_tail_emu.so")
    compile_emu("attn_tail_emu.cpp", so)
    lib = ctypes.CDLL(so)
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    lib.emu_attn_tail_fwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, ctypes.c_uint32, ctypes.c_float, ctypes.c_uint64]
    lib.emu_attn_tail_bwd.argtypes = [vp] * 6 + [i32, i32, i32, i32, ctypes.c_uint32, ctypes.c_float, ctypes.c_uint64]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_tail_rows_rule(emu):
    assert [emu.emu_attn_tail_rows(T) for T in (64, 128, 129, 131, 144, 145, 200, 256, 644, 1156, 1152)] == [0, 0, 1, 3, 16, 0, 0, 0, 4, 4, 0]
    assert emu.emu_attn_tail_rows(128 * 40 + 16) == 0          # score rows would not fit in shared memory: no split


@pytest.mark.parametrize("B,T,H,p", [(2, 131, 2, 0.0), (1, 260, 2, 0.1), (1, 137, 1, 0.1), (1, 144, 1, 0.25)])
def test_tail_forward_backward_vs_torch(emu, B, T, H, p):
    d = H * 64
    r = emu.emu_attn_tail_rows(T)
    assert r == T % 128
    Tm = T - r
    seed = 2 ** 41 + 5
    thresh16 = O.attn_dropout_thresh16(p) if p else 0
    keep_scale = 1.0 / (1.0 - thresh16 / 65536.0)
    torch.manual_seed(T)
    qkv = (torch.randn(B * T, 3 * d) * 0.7).bfloat16()
    dout = (torch.randn(B * T, d) * 0.5).bfloat16()
    mask = (torch.from_numpy(O.attn_dropout_keep_mask(seed, np.arange(B * H * T), T, p).reshape(B, H, T, T)).double() if p
            else torch.ones(B, H, T, T, dtype=torch.float64))
    q, k, v = [t.view(B, T, H, 64).transpose(1, 2).double().requires_grad_(True) for t in qkv.double().split(d, dim=1)]
    att = (q @ k.transpose(-1, -2)) * 0.125
    att = att.masked_fill(~torch.ones(T, T, dtype=torch.bool).tril(), float("-inf"))
    pm = torch.softmax(att, -1) * mask * keep_scale
    o_full = pm @ v                                                       # [B, H, T, 64]
    ref = o_full.transpose(1, 2).reshape(B * T, d)
    lse_ref = torch.logsumexp(att, -1)                                    # [B, H, T]
    rows = torch.arange(B * T).view(B, T)
    tail_rows, main_rows = rows[:, Tm:].reshape(-1), rows[:, :Tm].reshape(-1)

    # ---- forward: the tail rows of out / lse, nothing else touched
    out = torch.full((B * T, d), 7.0).bfloat16()
    lse = torch.full((B * H * T,), 3.0)
    rc = emu.emu_attn_tail_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, Tm, thresh16, keep_scale, seed)
    assert rc == 0, emu.emu_last_error()
    assert rel(out[tail_rows].float(), ref[tail_rows]) < 5e-3
    assert torch.all(out[main_rows].float() == 7.0)
    lse_v = lse.view(B, H, T)
    assert rel(lse_v[:, :, Tm:], lse_ref[:, :, Tm:]) < 1e-5
    assert torch.all(lse_v[:, :, :Tm] == 3.0)

    # ---- backward: gradient of the WHOLE attention; the tile kernel's share (query rows < Tm) is put in place as bf16 first
    ref.backward(dout.double())
    dref = torch.cat([t.grad.transpose(1, 2).reshape(B * T, d) for t in (q, k, v)], dim=1)             # [B*T, 3d]
    q2, k2, v2 = [t.view(B, T, H, 64).transpose(1, 2).double().requires_grad_(True) for t in qkv.double().split(d, dim=1)]
    att2 = (q2 @ k2.transpose(-1, -2)) * 0.125
    att2 = att2.masked_fill(~torch.ones(T, T, dtype=torch.bool).tril(), float("-inf"))
    o2 = (torch.softmax(att2, -1) * mask * keep_scale) @ v2
    go = dout.double().view(B, T, H, 64).transpose(1, 2).clone()
    go[:, :, Tm:] = 0                                                     # only the query rows below Tm
    o2.backward(go)
    dmain = torch.cat([t.grad.transpose(1, 2).reshape(B * T, d) for t in (q2, k2, v2)], dim=1)
    dqkv = dmain.bfloat16().contiguous()
    dqkv[tail_rows, d:] = 9.0                                             # dK / dV rows of the tail keys: never written by the tile kernel
    dqkv[:, :d] = 5.0                                                     # dQ goes through dq_acc, not through dqkv
    delta = (o_full.detach() * dout.double().view(B, T, H, 64).transpose(1, 2)).sum(-1).reshape(-1).float().contiguous()      # [B*H*T]
    lse_in = lse_ref.reshape(-1).float().contiguous()
    dq_acc = torch.full((B * T, d), 11.0)
    rc = emu.emu_attn_tail_bwd(qkv.data_ptr(), dout.data_ptr(), lse_in.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), dq_acc.data_ptr(), B, T, H, Tm,
                               thresh16, keep_scale, seed)
    assert rc == 0, emu.emu_last_error()
    assert torch.all(dqkv[:, :d].float() == 5.0) and torch.all(dq_acc[main_rows] == 11.0)
    # dQ of the tail rows: dq_acc holds sum_j dS~ k_j, attn_dq_convert multiplies by scale / (1 - p)
    assert rel(dq_acc[tail_rows] * 0.125 * keep_scale, dref[tail_rows, :d]) < 8e-3
    assert rel(dqkv[:, d:2 * d].float(), dref[:, d:2 * d]) < 8e-3
    assert rel(dqkv[:, 2 * d:].float(), dref[:, 2 * d:]) < 8e-3
    # and the tail rows really matter: without their share the K / V gradients are off by far more than the tolerance
    assert rel(dmain[:, d:2 * d], dref[:, d:2 * d]) > 3e-2
