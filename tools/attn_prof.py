"""Attention kernels alone at the cfg3 shape (B=32, H=16, T=1156): time fwd / bwd with and without dropout.
TTTS_ATTN_LEGACY=1 selects the mma.sync kernels.  Target for `ncu --set full -k regex:attn_`."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ttts_b200 import _lib as L
from ttts_b200.gpt import engine as E

lib = L.lib(); E._setup_prototypes(lib)
B, T, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 1156, 16)))
iters = int(os.environ.get("ITERS", "5"))
d = H * 64
torch.manual_seed(0)
qkv = (torch.randn(B * T, 3 * d, device="cuda") * 0.7).bfloat16()
out = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
dout = (torch.randn(B * T, d, device="cuda") * 0.5).bfloat16()
lse = torch.zeros(B * H * T, device="cuda")
dqkv = torch.zeros_like(qkv)
scratch = torch.zeros(B * H * T + 64 + B * T * d, device="cuda")
flops_fwd = 4.0 * B * H * T * T * 64 / 2


def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for p in ([float(os.environ['ONLY_P'])] if 'ONLY_P' in os.environ else [0.0, 0.1]):
    f = lambda: L.check(lib.ttts_attn_fwd(L.ptr(qkv), L.ptr(out), L.ptr(lse), B, T, H, ctypes.c_float(p), ctypes.c_uint64(7), L.stream_ptr()))
    b = lambda: L.check(lib.ttts_attn_bwd(L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(scratch), L.ptr(dqkv), B, T, H, ctypes.c_float(p),
                                          ctypes.c_uint64(7), L.stream_ptr()))
    mf, mb = t(f), t(b)
    print("legacy=%s p=%.1f  fwd %.3f ms (%.0f TFLOP/s)   bwd %.3f ms (%.0f TFLOP/s, 2.5x fwd flops)" % (
        os.environ.get("TTTS_ATTN_LEGACY", "0"), p, mf, flops_fwd / mf / 1e9, mb, 2.5 * flops_fwd / mb / 1e9), flush=True)
    # digests of the results: runs with different TTTS_ATTN_VER must print identical lines (the versions differ in schedule only)
    import hashlib
    torch.cuda.synchronize()
    dig = [hashlib.sha1(x.detach().cpu().view(torch.uint8).numpy().tobytes()).hexdigest()[:12] for x in (out, lse, dqkv)]
    print("digest p=%.1f out %s lse %s dqkv %s" % (p, *dig), flush=True)
