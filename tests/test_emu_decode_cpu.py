"""CPU emulation of the KV-cache decode kernels: tests/emu builds ttts_b200/csrc/gpt_decode.cu -- the same source nvcc compiles, kernels AND
launch sequence -- for the host (one OS thread per CUDA thread, real barriers) and this test runs it against the oracle's cached decode
(oracle/gpt_oracle.py, pinned to the REAL reference by tests/test_oracle_kv_decode.py).  It checks what can be checked without a GPU: index
arithmetic, cache layout, partial-sum plumbing, barrier placement, the device-resident slot counter.  It is NOT the parity test of the
product (that is tests/test_gpu_gpt.py::test_kv_decode_step_matches_oracle on a B200)."""
import ctypes
import os
import shutil
import subprocess
import sys

import pytest
import torch

from oracle import gpt_oracle as O
from ttts_b200.gpt.engine import GptConfig, GptDecode, tensor_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu_build import compile_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libdecode_emu.so")
    # TTTS_EMU_CXXFLAGS="-g -fsanitize=address" (with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0) turns every
    # out-of-bounds shared / global access of the emulated kernels into a hard error
    compile_emu("decode_emu.cpp", so)
    lib = ctypes.CDLL(so)
    lib.emu_param_off.restype = ctypes.c_longlong
    lib.emu_param_off.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int, ctypes.c_int]
    lib.emu_param_count.restype = ctypes.c_longlong
    lib.emu_param_count.argtypes = [ctypes.POINTER(GptConfig)]
    lib.emu_kv_bytes.restype = ctypes.c_longlong
    lib.emu_kv_bytes.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int, ctypes.c_int]
    lib.emu_decode_workspace_bytes.restype = ctypes.c_longlong
    lib.emu_decode_workspace_bytes.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int]
    lib.emu_gpt_decode_step.argtypes = [ctypes.POINTER(GptDecode)]
    lib.emu_kv_fill_layer.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.emu_last_error.restype = ctypes.c_char_p
    lib.emu_launches.restype = ctypes.c_ulonglong
    return lib


def _gpt_config(cfg):
    c = GptConfig()
    c.layers, c.model_dim, c.heads = cfg["layers"], cfg["model_dim"], cfg["heads"]
    c.max_text_tokens, c.max_mel_tokens = cfg["max_text_tokens"], cfg["max_mel_tokens"]
    c.n_text_vocab, c.n_mel_vocab = cfg["number_text_tokens"] + 1, cfg["number_mel_codes"]
    c.start_text_token, c.stop_text_token = cfg["start_text_token"], 0
    c.start_mel_token, c.stop_mel_token = cfg["start_mel_token"], cfg["stop_mel_token"]
    c.mel_length_compression = 1024
    return c


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.mark.parametrize("dims,B,pos_shift", [((2, 128, 2), 3, 0), ((2, 128, 2), 5, 1), ((1, 256, 4), 2, 0)])
def test_decode_kernels_on_the_cpu_emulation(emu, dims, B, pos_shift):
    layers, d, heads = dims
    cfg = O.default_config(layers=layers, model_dim=d, heads=heads, max_text_tokens=20, max_mel_tokens=30)
    params = O.init_params(cfg, seed=1)
    # the oracle init keeps biases / position tables small; give every tensor the decode path reads a visible scale so that a wrong
    # offset, a dropped bias or a shifted position row cannot hide under the tolerance
    g = torch.Generator().manual_seed(3)
    for k in params:
        if k.endswith(".bias") or "pos_embedding" in k:
            params[k] = params[k] + (0.3 if k.endswith(".bias") else 0.05) * torch.randn(params[k].shape, generator=g)
    c = _gpt_config(cfg)
    total = emu.emu_param_count(ctypes.byref(c))
    assert total > 0
    flat = torch.zeros(total, dtype=torch.float32)
    for name, tid, layer, shape in tensor_table(c):
        off = emu.emu_param_off(ctypes.byref(c), tid, layer)
        assert off >= 0
        flat[off:off + params[name].numel()] = params[name].reshape(-1)
    flat16 = flat.to(torch.bfloat16)

    TL, mc, steps = 5, 3, 3
    text = torch.randint(1, 255, (B, TL), generator=g)
    codes = torch.randint(0, 1024, (B, mc + steps + 1), generator=g)
    text_in = torch.cat([torch.full((B, 1), cfg["start_text_token"]), text, torch.zeros(B, 1, dtype=torch.int64)], 1)
    mel_in = torch.cat([torch.full((B, 1), cfg["start_mel_token"]), codes[:, :mc]], 1)
    Tt, T = TL + 2, TL + 2 + mc + 1
    T_max = 64
    with torch.no_grad():
        cache, slot, _ = O.kv_prefill(params, cfg, text_in, mel_in, T_max=T_max, emulate_bf16=True)
        cache32, slot32, _ = O.kv_prefill(params, cfg, text_in, mel_in, T_max=T_max)
    assert slot == T

    # ---- cache fill kernel: packed c_attn output [B*T, 3d] of every layer -> [L, 2, B, H, T_max, 64] ----
    kv = torch.zeros(emu.emu_kv_bytes(ctypes.byref(c), B, T_max) // 2, dtype=torch.bfloat16)
    kv6 = kv.view(layers, 2, B, heads, T_max, 64)
    for l in range(layers):
        k = cache[l, 0, :, :, :T].permute(0, 2, 1, 3).reshape(B * T, d)            # [B, H, T, 64] -> rows (b, t), columns (h, j)
        v = cache[l, 1, :, :, :T].permute(0, 2, 1, 3).reshape(B * T, d)
        qkv = torch.cat([torch.full_like(k, 7.0), k, v], 1).to(torch.bfloat16).contiguous()
        rc = emu.emu_kv_fill_layer(qkv.data_ptr(), B, T, d, heads, T, kv6[l, 0].data_ptr(), kv6[l, 1].data_ptr(), T_max)
        assert rc == 0, emu.emu_last_error()
    assert torch.equal(kv6[:, :, :, :, :T].float(), cache[:, :, :, :, :T])
    assert float(kv6[:, :, :, :, T:].float().abs().max()) == 0.0

    # ---- decode steps ----
    ws = torch.zeros(emu.emu_decode_workspace_bytes(ctypes.byref(c), B) + 256, dtype=torch.uint8)
    ws_off = (-ws.data_ptr()) % 256
    slot_t = torch.tensor([slot], dtype=torch.int32)
    logits = torch.zeros(B, cfg["number_mel_codes"], dtype=torch.float32)
    a = GptDecode()
    a.cfg = c
    a.B, a.T_max, a.text_positions, a.pos_shift = B, T_max, Tt, pos_shift
    a.codes, a.ld_codes = codes.data_ptr(), codes.stride(0)
    a.slot = slot_t.data_ptr()
    a.params, a.params16 = flat.data_ptr(), flat16.data_ptr()
    a.kv, a.kv_bytes = kv.data_ptr(), kv.numel() * 2
    a.workspace, a.workspace_bytes = ws.data_ptr() + ws_off, ws.numel() - 256
    a.logits = logits.data_ptr()
    launches0 = emu.emu_launches()
    for s in range(steps):
        n = mc + s + 1
        with torch.no_grad():
            want = O.kv_decode_step(params, cfg, cache, slot, codes[:, n - 1], Tt, pos_shift=pos_shift, emulate_bf16=True)
            want32 = O.kv_decode_step(params, cfg, cache32, slot32, codes[:, n - 1], Tt, pos_shift=pos_shift)
        slot += 1
        slot32 += 1
        rc = emu.emu_gpt_decode_step(ctypes.byref(a))
        assert rc == 0, emu.emu_last_error()
        assert int(slot_t[0]) == slot                                    # advanced on the "device"
        got = logits.clone()
        assert torch.isfinite(got).all()
        # same roundings as the bf16-emulating oracle up to summation order and single-vs-double rounding of (acc + bias)
        assert rel(got, want) <= 6e-3, (s, rel(got, want))
        assert rel(got, want32) <= 2e-2, (s, rel(got, want32))             # the stated GPU tolerance against the fp32 oracle
        # the step appended this token's K / V rows for every layer
        assert rel(kv6[:, :, :, :, slot - 1].float(), cache[:, :, :, :, slot - 1]) <= 6e-3
    # fused schedule (default for B <= 4): 5 launches per layer + the embedding row-op + final row-op, head, advance; TTTS_DECODE_FUSE=0: 8 per layer + 3
    fused = {"0": False, "1": True}.get(os.environ.get("TTTS_DECODE_FUSE"), B <= 4)
    per_step = 5 * layers + 4 if fused else 8 * layers + 3
    assert emu.emu_launches() - launches0 == steps * per_step
    # capacity: a full cache refuses to write (the host wrapper raises before this; the kernels must stay in bounds)
    slot_t[0] = T_max
    before = kv.clone()
    assert emu.emu_gpt_decode_step(ctypes.byref(a)) == 0
    assert torch.equal(before.view(torch.int16), kv.view(torch.int16))
    # bad arguments are reported, not executed
    a.pos_shift = 2
    assert emu.emu_gpt_decode_step(ctypes.byref(a)) != 0 and b"pos_shift" in emu.emu_last_error()
