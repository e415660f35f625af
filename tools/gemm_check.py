"""GPU bring-up check for the tcgen05 GEMM: all operand-major combos, ragged shapes, epilogues, split-K.
Run on a B200:  python tools/gemm_check.py  (prints PASS/FAIL per case; exits non-zero on failure)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ttts_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
fails = 0


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def run(M, N, K, a_mn, b_mn, epi=L.EPI_F32, split_k=1, ldpad=0):
    global fails
    Af = torch.randn(M, K, device=dev) * 0.5
    Bf = torch.randn(K, N, device=dev) * 0.5
    A16 = Af.bfloat16(); B16 = Bf.bfloat16()
    ref = A16.float() @ B16.float()
    def padded(t):   # TMA needs 16B-aligned row strides: pad the leading dimension to a multiple of 64 elements
        r, c = t.shape
        cp = (c + 63) // 64 * 64
        buf = torch.full((r, cp), 7.0, device=dev, dtype=t.dtype)
        buf[:, :c] = t
        return buf[:, :c]
    A_in = padded(A16.t()) if a_mn else padded(A16)          # [K,M] or [M,K]
    B_in = padded(B16) if b_mn else padded(B16.t())          # [K,N] or [N,K]
    bias = torch.randn(N, device=dev)
    ldo = N + ldpad
    name = "M%d N%d K%d a_mn%d b_mn%d epi%d split%d" % (M, N, K, a_mn, b_mn, epi, split_k)
    if epi == L.EPI_F32:
        out = torch.full((M, ldo), 777.0, device=dev)
        L.gemm(A_in, B_in, out, a_mn=a_mn, b_mn=b_mn, epi=epi, bias=bias, M=M, N=N, K=K)
        got = out[:, :N]; want = ref + bias
        ok_pad = bool((out[:, N:] == 777.0).all())
    elif epi == L.EPI_BF16:
        out = torch.full((M, ldo), 777.0, device=dev, dtype=torch.bfloat16)
        L.gemm(A_in, B_in, out, a_mn=a_mn, b_mn=b_mn, epi=epi, bias=bias, M=M, N=N, K=K)
        got = out[:, :N]; want = (ref + bias)
        ok_pad = bool((out[:, N:] == 777.0).all())
    elif epi == L.EPI_F32_ADD:
        out = torch.ones((M, ldo), device=dev)
        L.gemm(A_in, B_in, out, a_mn=a_mn, b_mn=b_mn, epi=epi, split_k=split_k, M=M, N=N, K=K)
        got = out[:, :N]; want = ref + 1.0
        ok_pad = bool((out[:, N:] == 1.0).all())
    elif epi == L.EPI_GELU:
        out = torch.zeros((M, ldo), device=dev, dtype=torch.bfloat16)
        pre = torch.zeros((M, ldo), device=dev, dtype=torch.bfloat16)
        L.gemm(A_in, B_in, out, a_mn=a_mn, b_mn=b_mn, epi=epi, bias=bias, aux_out=pre, M=M, N=N, K=K)
        want_pre = (ref + bias).bfloat16()
        want = torch.nn.functional.gelu(want_pre.float(), approximate="tanh")
        got = out[:, :N]
        ok_pad = rel(pre[:, :N], want_pre) < 1e-2
    elif epi == L.EPI_RESID:
        x = torch.randn(M, ldo, device=dev)
        out = torch.zeros((M, ldo), device=dev)
        L.gemm(A_in, B_in, out, a_mn=a_mn, b_mn=b_mn, epi=epi, bias=bias, aux=x, M=M, N=N, K=K)
        want = x[:, :N] + (ref + bias).bfloat16().float()
        got = out[:, :N]; ok_pad = True
    elif epi == L.EPI_DGELU:
        pre = (torch.randn(M, ldo, device=dev)).bfloat16()
        out = torch.zeros((M, ldo), device=dev, dtype=torch.bfloat16)
        L.gemm(A_in, B_in, out, a_mn=a_mn, b_mn=b_mn, epi=epi, aux=pre, M=M, N=N, K=K)
        p = pre[:, :N].float().requires_grad_(True)
        torch.nn.functional.gelu(p, approximate="tanh").backward(ref)
        want = p.grad; got = out[:, :N]; ok_pad = True
    torch.cuda.synchronize()
    r = rel(got, want)
    tol = 1e-2 if epi in (L.EPI_BF16, L.EPI_GELU, L.EPI_DGELU) else 2e-3
    ok = (r < tol) and ok_pad
    print("%s  %-60s rel=%.3e pad_ok=%s" % ("PASS" if ok else "FAIL", name, r, ok_pad), flush=True)
    if not ok:
        fails += 1
        d = (got.float() - want.float()).abs()
        print("   max abs diff %.4f at %s ; got[0,:4]=%s want[0,:4]=%s" % (d.max().item(), divmod(d.argmax().item(), N),
              got[0, :4].tolist(), want[0, :4].tolist()), flush=True)


if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0), "lib ok:", L.lib().ttts_device_ok())
    # smallest first: one tile, one k-block
    for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        run(128, 256, 64, a_mn, b_mn)
    for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        run(128, 128, 64, a_mn, b_mn)
        run(256, 512, 256, a_mn, b_mn)
        run(1000, 264, 328, a_mn, b_mn, ldpad=8)         # ragged everything (K tail: 328 = 5*64+8)
    run(5152, 1536, 512, 0, 1, epi=L.EPI_BF16)
    run(5152, 2048, 512, 0, 1, epi=L.EPI_GELU)
    run(5152, 512, 2048, 0, 1, epi=L.EPI_RESID)
    run(5152, 2048, 512, 0, 0, epi=L.EPI_DGELU)
    run(4112, 1026, 512, 0, 0, epi=L.EPI_BF16, ldpad=62)  # mel head, ldo=1088
    run(1040, 257, 512, 0, 0, epi=L.EPI_BF16, ldpad=63)   # text head
    run(512, 1536, 5152, 1, 1, epi=L.EPI_F32_ADD, split_k=3)
    run(1026, 512, 4112, 1, 1, epi=L.EPI_F32_ADD, split_k=4)
    run(4112, 512, 1026, 0, 1, epi=L.EPI_BF16)             # head dgrad: K=1026 ragged reduction
    # timing at cfg3 shapes
    for (M, N, K, a_mn, b_mn, epi, sk) in [(36992, 3072, 1024, 0, 1, L.EPI_BF16, 1), (36992, 1024, 4096, 0, 1, L.EPI_RESID, 1),
                                          (36992, 4096, 1024, 0, 1, L.EPI_GELU, 1), (36992, 1024, 4096, 0, 0, L.EPI_BF16, 1),
                                          (1024, 4096, 36992, 1, 1, L.EPI_F32_ADD, 8), (1024, 1024, 36992, 1, 1, L.EPI_F32_ADD, 9)]:
        A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
        B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
        out_dt = torch.float32 if epi in (L.EPI_RESID, L.EPI_F32_ADD) else torch.bfloat16
        out = torch.zeros(M, N, device=dev, dtype=out_dt)
        kw = dict(a_mn=a_mn, b_mn=b_mn, epi=epi, split_k=sk)
        if epi == L.EPI_RESID: kw["aux"] = out
        if epi == L.EPI_GELU: kw["aux_out"] = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3): L.gemm(A, B, out, **kw)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): L.gemm(A, B, out, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("TIME M%d N%d K%d a_mn%d b_mn%d epi%d split%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, a_mn, b_mn, epi, sk, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
        t0 = time.time()
        Af = A.t() if a_mn else A
        Bf = B if b_mn else B.t()
        for _ in range(3): torch.matmul(Af, Bf)
        e0.record()
        for _ in range(10): torch.matmul(Af, Bf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("     cuBLAS same shape: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9), flush=True)
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)
