"""CPU emulation of the tiling and index formulas of the experimental tensor-core convolution (ttts_b200/csrc/conv1d_tc.cu): per 128-frame
tile and 64-wide k-block it builds the im2col tile A[t, j] and the weight tile B[co, j] exactly as the kernel's worker threads do
(j = tt * C + ci, tap k = kb * TAPS + tt, zero columns past the last tap, zero-filled window outside the clip) and checks that the sum of
A B^T over the k-blocks equals F.conv1d -- with exact operands and with the split-bf16 products the kernel issues.  It does NOT model the
128B swizzle, the UMMA descriptors or the barriers; those mirror the attention kernels' P path and need the hardware.

    python tools/conv_tc_emulation.py
"""
import torch
import torch.nn.functional as F


def emulate(C, K, DIL, T=300, split=True):
    TAPS = 64 // C
    NKB = (K + TAPS - 1) // TAPS
    pad = DIL * (K - 1) // 2
    x = torch.randn(C, T)
    w = torch.randn(C, C, K) * 0.1
    xl = torch.where(x > 0, x, 0.1 * x)
    ref = F.conv1d(xl[None], w, padding=pad, dilation=DIL)[0]
    out = torch.zeros(C, T)
    bf = lambda v: v.to(torch.bfloat16).float()
    for t0 in range(0, T, 128):
        W = 128 + (K - 1) * DIL
        win = torch.zeros(C, W)
        for u in range(W):
            ti = t0 - pad + u
            if 0 <= ti < T:
                win[:, u] = xl[:, ti]
        acc = torch.zeros(128, C)
        for kb in range(NKB):
            A, Bt = torch.zeros(128, 64), torch.zeros(C, 64)
            for c16 in range(8):
                tt, ci0 = (c16 * 8) // C, (c16 * 8) % C
                k = kb * TAPS + tt
                for e in range(8):
                    if k < K:
                        A[:, c16 * 8 + e] = win[ci0 + e, k * DIL: k * DIL + 128]
                        Bt[:, c16 * 8 + e] = w[:, ci0 + e, k]
            if split:
                Ah, Bh = bf(A), bf(Bt)
                Al, Bl = bf(A - Ah), bf(Bt - Bh)
                acc += Ah @ Bh.T + Ah @ Bl.T + Al @ Bh.T
            else:
                acc += A @ Bt.T
        n = min(128, T - t0)
        out[:, t0:t0 + n] = acc[:n].T
    return float((out - ref).abs().max() / ref.abs().max())


if __name__ == "__main__":
    torch.manual_seed(0)
    for C in (32, 64):
        for K in (3, 7, 11):
            for D in (1, 3, 5):
                e0, e1 = emulate(C, K, D, split=False), emulate(C, K, D, split=True)
                assert e0 < 1e-5 and e1 < 1e-4, (C, K, D, e0, e1)
                print("C %2d K %2d DIL %d: exact operands %.1e, split bf16 %.1e" % (C, K, D, e0, e1))
