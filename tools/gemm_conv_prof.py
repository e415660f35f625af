"""One wide convolution of the training tapes on the split-bf16 GEMM route (ttts_b200/vqvae/train_encoder.py: `_gemm_conv_fwd / _gemm_conv_bwd`),
forward + backward, timed with CUDA events; the target of `ncu --set full -k regex:gemm` for the tap-concatenated GEMM (overlapped operand view).
    python tools/gemm_conv_prof.py [B Cin T Cout K stride dil pad]        default: a Generator ResBlock layer, 64 x 128 x 2560, K = 11"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ttts_b200.vqvae.train_encoder import CudaKernels

a = [int(v) for v in sys.argv[1:9]] if len(sys.argv) >= 9 else [64, 128, 2560, 128, 11, 1, 1, 5]
B, Cin, T, Cout, K, stride, dil, pad = a
iters = int(os.environ.get("ITERS", "5"))
Kn = CudaKernels()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, Cin, T, device="cuda", generator=g)
w = torch.randn(Cout, Cin, K, device="cuda", generator=g) / (Cin * K) ** 0.5
b = torch.randn(Cout, device="cuda", generator=g)
route = ("split-bf16 tcgen05 GEMM, tap_concat=%d" % int(Kn.tap_concat)) if Kn._gemm_ok(x, w, stride, dil, pad, 1) else "exact-fp32 CUDA-core kernels"
y = Kn.conv_fwd(x, w, b, stride, dil, pad, True)
dy = torch.randn_like(y)


def t(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


fl = 2.0 * B * y.shape[-1] * Cin * Cout * K
mf = t(lambda: Kn.conv_fwd(x, w, b, stride, dil, pad, True))
mb = t(lambda: Kn.conv_bwd(dy, x, w, stride, dil, pad, True, True, True))
print("%s:  B %d Cin %d T %d Cout %d K %d s %d d %d: forward %.3f ms (%.1f TFLOP/s fp32-equivalent), backward %.3f ms (%.1f)" % (
    route, B, Cin, T, Cout, K, stride, dil, mf, fl / mf / 1e9, mb, 2 * fl / mb / 1e9), flush=True)
