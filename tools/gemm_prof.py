"""One GEMM shape, a few launches -- target for `ncu --set full` (TTTS_GEMM_2CTA=1 selects the CTA-pair kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ttts_b200 import _lib as L

M, N, K = 36992, 3072, 1024
A = torch.randn(M, K, device="cuda").bfloat16()
B = torch.randn(K, N, device="cuda").bfloat16()
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(N, device="cuda")
for _ in range(4):
    L.gemm(A, B, out, b_mn=True, epi=L.EPI_BF16, bias=bias)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.gemm(A, B, out, b_mn=True, epi=L.EPI_BF16, bias=bias)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("%.3f ms %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
