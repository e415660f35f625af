"""Every GEMM of one cfg3 transformer layer (forward + backward), exactly as gpt_engine.cu launches it (operand majors, epilogue,
split-K), timed alone with CUDA events: the per-shape table behind the step's `roofline` line.  Under
`ncu --set full -k regex:gemm2 ...` pass `ONCE=1` to launch each shape a single time (`ONLY=name[,name]` selects shapes).

    python tools/gemm_step_prof.py            # table
    ONCE=1 ONLY=fc1_gelu,proj_resid ncu --set full --import-source on --clock-control none -k regex:gemm2 -o gpurun_out/x python tools/gemm_step_prof.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ttts_b200 import _lib as L

M, d = 36992, 1024
once = os.environ.get("ONCE") == "1"
only = [s for s in os.environ.get("ONLY", "").split(",") if s]
iters = 1 if once else int(os.environ.get("ITERS", "10"))
dev = "cuda"
torch.manual_seed(0)


def bf(*shape, s=0.5):
    return (torch.randn(*shape, device=dev) * s).bfloat16()


h = bf(M, d)                     # ln output / attention output / bf16 gradient of a [M, d] tensor
big = bf(M, 4 * d)               # gelu output / d(pre)
qkvg = bf(M, 3 * d)              # d(qkv)
W_qkv, W_proj, W_fc, W_pr = bf(d, 3 * d, s=0.02), bf(d, d, s=0.02), bf(d, 4 * d, s=0.02), bf(4 * d, d, s=0.02)   # HF Conv1D layout [in, out]
x32 = torch.randn(M, d, device=dev)
o32 = torch.empty(M, d, device=dev)
o_qkv = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
o_big = torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16)
o_pre = torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16)
o_h = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
g_qkv, g_proj, g_fc, g_pr = (torch.zeros_like(w, dtype=torch.float32) for w in (W_qkv, W_proj, W_fc, W_pr))
b_d, b_3d, b_4d = torch.randn(d, device=dev), torch.randn(3 * d, device=dev), torch.randn(4 * d, device=dev)


def pick_split_k(Mo, No, Ko, clusters=74):
    """mirror of pick_split_k2 (gemm2_tcgen05.cu) for the CTA-pair kernel: 256 x 256 tiles, 64-wide k-blocks"""
    tiles = ((Mo + 255) // 256) * ((No + 255) // 256)
    kblocks = (Ko + 63) // 64
    best, best_eff = 1, -1.0
    for s in range(1, min(32, kblocks) + 1):
        if kblocks // s < 8 and s > 1:
            break
        items = tiles * s
        waves = (items + clusters - 1) // clusters
        eff = items / (waves * clusters)
        if eff > best_eff + 0.02:
            best_eff, best = eff, s
    return best


cases = {
    # forward (A K-major, B = weight consumed in place as MN-major)
    "qkv_bf16": (M, 3 * d, d, lambda: L.gemm(h, W_qkv, o_qkv, b_mn=True, epi=L.EPI_BF16, bias=b_3d)),
    "proj_resid": (M, d, d, lambda: L.gemm(h, W_proj, o32, b_mn=True, epi=L.EPI_RESID, bias=b_d, aux=x32, drop_p=0.1, drop_seed=3)),
    "fc1_gelu": (M, 4 * d, d, lambda: L.gemm(h, W_fc, o_big, b_mn=True, epi=L.EPI_GELU, bias=b_4d, aux_out=o_pre)),
    "fc2_resid": (M, d, 4 * d, lambda: L.gemm(big, W_pr, o32, b_mn=True, epi=L.EPI_RESID, bias=b_d, aux=x32, drop_p=0.1, drop_seed=4)),
    # backward dgrad (B = the same weight read as K-major [N=in, K=out])
    "dgrad_pr_dgelu": (M, 4 * d, d, lambda: L.gemm(h, W_pr, o_big, epi=L.EPI_DGELU, aux=o_pre)),
    "dgrad_fc": (M, d, 4 * d, lambda: L.gemm(big, W_fc, o_h, epi=L.EPI_BF16)),
    "dgrad_proj": (M, d, d, lambda: L.gemm(h, W_proj, o_h, epi=L.EPI_BF16)),
    "dgrad_qkv": (M, d, 3 * d, lambda: L.gemm(qkvg, W_qkv, o_h, epi=L.EPI_BF16)),
}


def wgrad(name, X, dY, G):
    Mo, No, Ko = X.shape[1], dY.shape[1], X.shape[0]
    sk = pick_split_k(Mo, No, Ko)
    cases[name] = (Mo, No, Ko, lambda: L.gemm(X, dY, G, a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=sk))


wgrad("wgrad_pr", big, h, g_pr)
wgrad("wgrad_fc", h, big, g_fc)
wgrad("wgrad_proj", h, h, g_proj)
wgrad("wgrad_qkv", h, qkvg, g_qkv)

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot_ms = tot_fl = 0.0
for name, (Mo, No, Ko, fn) in cases.items():
    if only and name not in only:
        continue
    if not once:
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()                                   # operands come from HBM, as inside the step
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    ms /= iters
    fl = 2.0 * Mo * No * Ko
    tot_ms += ms; tot_fl += fl
    print("%-16s M %6d N %5d K %6d  %8.1f us  %7.1f TFLOP/s" % (name, Mo, No, Ko, ms * 1e3, fl / ms / 1e9), flush=True)
if tot_ms:
    print("layer total %.3f ms, %.1f TFLOP/s" % (tot_ms, tot_fl / tot_ms / 1e9))
