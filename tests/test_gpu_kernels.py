"""GPU (-m gpu): per-kernel parity of the sm_100a kernels, called through the C ABI, against plain fp32 torch math."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from ttts_b200 import _lib
    from ttts_b200.gpt import engine as E
    E._setup_prototypes(_lib.lib())
    assert _lib.lib().ttts_device_ok() == 1
    return _lib


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def _padded(t):
    r, c = t.shape
    cp = (c + 63) // 64 * 64
    buf = torch.full((r, cp), 7.0, device=t.device, dtype=t.dtype)
    buf[:, :c] = t
    return buf[:, :c]


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (1000, 264, 328), (130, 128, 1026)])
def test_gemm_majors_and_ragged(L, a_mn, b_mn, M, N, K):
    torch.manual_seed(0)
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(K, N, device="cuda") * 0.5).bfloat16()
    ref = A.float() @ B.float()
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N + 8), 777.0, device="cuda")
    L.gemm(_padded(A.t()) if a_mn else _padded(A), _padded(B) if b_mn else _padded(B.t()), out, a_mn=a_mn, b_mn=b_mn, epi=L.EPI_F32, bias=bias,
           M=M, N=N, K=K)
    assert rel(out[:, :N], ref + bias) < 1e-5
    assert bool((out[:, N:] == 777.0).all())            # nothing written outside the logical tile


def test_gemm_epilogues(L):
    torch.manual_seed(1)
    M, N, K = 644, 512, 256
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(K, N, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.1
    ref = A.float() @ W.float() + bias
    o16 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(A, W, o16, b_mn=True, epi=L.EPI_BF16, bias=bias)
    assert torch.equal(o16, ref.bfloat16()) or rel(o16, ref) < 3e-3
    pre = torch.zeros_like(o16)
    L.gemm(A, W, o16, b_mn=True, epi=L.EPI_GELU, bias=bias, aux_out=pre)
    want = torch.nn.functional.gelu(ref.bfloat16().float(), approximate="tanh")
    assert rel(o16, want) < 5e-3 and rel(pre, ref) < 3e-3
    x = torch.randn(M, N, device="cuda")
    o32 = torch.zeros(M, N, device="cuda")
    L.gemm(A, W, o32, b_mn=True, epi=L.EPI_RESID, bias=bias, aux=x)
    assert rel(o32, x + ref.bfloat16().float()) < 1e-3
    # dropout in the residual epilogue: keep-rate, scaling, determinism per seed
    o_a = torch.zeros(M, N, device="cuda"); o_b = torch.zeros(M, N, device="cuda"); o_c = torch.zeros(M, N, device="cuda")
    L.gemm(A, W, o_a, b_mn=True, epi=L.EPI_RESID, bias=bias, aux=torch.zeros_like(x), drop_p=0.1, drop_seed=5)
    L.gemm(A, W, o_b, b_mn=True, epi=L.EPI_RESID, bias=bias, aux=torch.zeros_like(x), drop_p=0.1, drop_seed=5)
    L.gemm(A, W, o_c, b_mn=True, epi=L.EPI_RESID, bias=bias, aux=torch.zeros_like(x), drop_p=0.1, drop_seed=6)
    assert torch.equal(o_a, o_b) and not torch.equal(o_a, o_c)
    full = ref.bfloat16().float()
    kept = o_a != 0
    assert abs(kept.float().mean().item() - 0.9) < 0.01
    assert rel(o_a[kept], full[kept] / 0.9) < 2e-3
    # split-K weight gradient accumulates into the existing buffer
    X = (torch.randn(2000, 256, device="cuda") * 0.5).bfloat16()
    dY = (torch.randn(2000, 384, device="cuda") * 0.5).bfloat16()
    dW = torch.ones(256, 384, device="cuda")
    L.gemm(X, dY, dW, a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=4)
    assert rel(dW, X.float().t() @ dY.float() + 1.0) < 1e-5


def test_gemm_rejects_bad_arguments(L):
    A = torch.zeros(128, 66, device="cuda", dtype=torch.bfloat16)      # row stride 132 B: not TMA-legal
    B = torch.zeros(128, 66, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(128, 128, device="cuda")
    with pytest.raises(L.TTTSError):
        L.gemm(A, B, out, epi=L.EPI_F32)
    with pytest.raises(L.TTTSError):
        L.gemm(torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16), torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16), out,
               epi=L.EPI_BF16, split_k=2)


@pytest.mark.parametrize("B,T,H", [(2, 64, 2), (2, 200, 2), (1, 644, 8), (3, 131, 1)])
def test_attention_fwd_bwd(L, B, T, H):
    lib = L.lib()
    d = H * 64
    torch.manual_seed(1)
    qkv = (torch.randn(B * T, 3 * d, device="cuda") * 0.7).bfloat16()
    out = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B * H * T, device="cuda")
    L.check(lib.ttts_attn_fwd(L.ptr(qkv), L.ptr(out), L.ptr(lse), B, T, H, ctypes.c_float(0.0), ctypes.c_uint64(0), L.stream_ptr()))
    q, k, v = [t.view(B, T, H, 64).transpose(1, 2).float().requires_grad_(True) for t in qkv.float().split(d, dim=1)]
    att = (q @ k.transpose(-1, -2)) * 0.125
    att = att.masked_fill(~torch.ones(T, T, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    ref = (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(B * T, d)
    assert rel(out, ref) < 5e-3
    assert rel(lse.view(B, H, T), torch.logsumexp(att, -1)) < 1e-5
    dout = (torch.randn(B * T, d, device="cuda") * 0.5).bfloat16()
    ref.backward(dout.float())
    dref = torch.cat([t.grad.transpose(1, 2).reshape(B * T, d) for t in (q, k, v)], dim=1)
    dqkv = torch.full_like(qkv, 5.0)
    delta = torch.zeros(B * H * T + 64 + B * T * d, device="cuda")
    L.check(lib.ttts_attn_bwd(L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(delta), L.ptr(dqkv), B, T, H, ctypes.c_float(0.0),
                              ctypes.c_uint64(0), L.stream_ptr()))
    for i in range(3):
        assert rel(dqkv[:, i * d:(i + 1) * d], dref[:, i * d:(i + 1) * d]) < 8e-3


def test_attention_tail_split_opt_in():
    """TTTS_ATTN_TAIL=1 (csrc/attention_tail.cu: the last T mod 128 <= 16 query rows on CUDA cores, the tile kernels stop at the last full
    tile): the same parity tests, in a child process (the switch is read once per process).  T = 644, 131, 389 take the split."""
    import subprocess, sys, os
    if os.environ.get("TTTS_TEST_CHILD") == "1":
        pytest.skip("child process of this very test")
    env = dict(os.environ, TTTS_ATTN_TAIL="1", TTTS_TEST_CHILD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__, "-k",
                        "test_attention_fwd_bwd or (test_attention_dropout_exact_with_mask and not 1-389-4-1 and not 2-200-2-1)"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("legacy", [0, 1])
@pytest.mark.parametrize("B,T,H", [(2, 200, 2), (1, 389, 4)])
def test_attention_dropout_exact_with_mask(L, B, T, H, legacy):
    """Attention-probability dropout, exactly: the device keep mask equals the oracle's numpy restatement of the hash bit for bit, and
    with that mask plain torch reproduces forward and backward (HF: modeling_gpt2.py:207-222: softmax -> dropout -> @ V)."""
    import subprocess, sys, os
    if legacy:      # the mma.sync kernels are selected per process (env read once): run this case in a child process
        env = dict(os.environ, TTTS_ATTN_LEGACY="1", TTTS_TEST_CHILD="1")
        if os.environ.get("TTTS_TEST_CHILD") != "1":
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu",
                                "%s::test_attention_dropout_exact_with_mask[%d-%d-%d-1]" % (__file__, B, T, H)], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            return
    import numpy as np
    from oracle import gpt_oracle as O
    lib = L.lib()
    lib.ttts_attn_dropout_mask.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p]
    lib.ttts_attn_dropout_mask.restype = ctypes.c_int
    d = H * 64
    p, seed = 0.1, 2 ** 40 + 17
    torch.manual_seed(5)
    qkv = (torch.randn(B * T, 3 * d, device="cuda") * 0.7).bfloat16()
    mask = torch.zeros(B * H, T, T, device="cuda", dtype=torch.uint8)
    L.check(lib.ttts_attn_dropout_mask(L.ptr(mask), B * H, T, ctypes.c_float(p), ctypes.c_uint64(seed), L.stream_ptr()))
    want = O.attn_dropout_keep_mask(seed, np.arange(B * H * T), T, p).reshape(B * H, T, T)
    assert np.array_equal(mask.cpu().numpy().astype(bool), want)
    keep_scale = 1.0 / (1.0 - O.attn_dropout_thresh16(p) / 65536.0)
    out = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B * H * T, device="cuda")
    L.check(lib.ttts_attn_fwd(L.ptr(qkv), L.ptr(out), L.ptr(lse), B, T, H, ctypes.c_float(p), ctypes.c_uint64(seed), L.stream_ptr()))
    q, k, v = [t.view(B, T, H, 64).transpose(1, 2).float().requires_grad_(True) for t in qkv.float().split(d, dim=1)]
    att = (q @ k.transpose(-1, -2)) * 0.125
    att = att.masked_fill(~torch.ones(T, T, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    pm = torch.softmax(att, -1) * mask.view(B, H, T, T).float() * keep_scale
    ref = (pm @ v).transpose(1, 2).reshape(B * T, d)
    assert rel(out, ref) < 6e-3
    dout = (torch.randn(B * T, d, device="cuda") * 0.5).bfloat16()
    ref.backward(dout.float())
    dref = torch.cat([t.grad.transpose(1, 2).reshape(B * T, d) for t in (q, k, v)], dim=1)
    dqkv = torch.full_like(qkv, 5.0)
    delta = torch.zeros(B * H * T + 64 + B * T * d, device="cuda")
    L.check(lib.ttts_attn_bwd(L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(delta), L.ptr(dqkv), B, T, H, ctypes.c_float(p),
                              ctypes.c_uint64(seed), L.stream_ptr()))
    for i in range(3):
        assert rel(dqkv[:, i * d:(i + 1) * d], dref[:, i * d:(i + 1) * d]) < 1e-2


def test_attention_dropout_statistics(L):
    """attn-prob dropout cannot be bit-matched to torch's Philox stream; check it is unbiased, seed-deterministic, and that
    backward uses the forward's mask (directional derivative vs finite difference of the SAME masked function)."""
    lib = L.lib()
    B, T, H = 2, 128, 2
    d = H * 64
    torch.manual_seed(3)
    qkv = (torch.randn(B * T, 3 * d, device="cuda") * 0.5).bfloat16()

    def fwd(x, p, seed):
        o = torch.zeros(B * T, d, device="cuda", dtype=torch.bfloat16)
        lse = torch.zeros(B * H * T, device="cuda")
        L.check(lib.ttts_attn_fwd(L.ptr(x), L.ptr(o), L.ptr(lse), B, T, H, ctypes.c_float(p), ctypes.c_uint64(seed), L.stream_ptr()))
        return o, lse
    o0, _ = fwd(qkv, 0.0, 0)
    o1, lse1 = fwd(qkv, 0.1, 11)
    o1b, _ = fwd(qkv, 0.1, 11)
    o2, _ = fwd(qkv, 0.1, 12)
    assert torch.equal(o1, o1b) and not torch.equal(o1, o2)
    acc = torch.zeros_like(o0, dtype=torch.float32)
    n = 64
    for s in range(n):
        acc += fwd(qkv, 0.1, 100 + s)[0].float()
    late = slice(B * T // 2, None)      # rows with many keys: the mean over masks converges to the undropped output
    assert rel((acc / n)[late], o0[late]) < 0.05
    # backward consistency with the same mask
    dout = (torch.randn(B * T, d, device="cuda") * 0.5).bfloat16()
    dqkv = torch.zeros_like(qkv); delta = torch.zeros(B * H * T + 64 + B * T * d, device="cuda")
    L.check(lib.ttts_attn_bwd(L.ptr(qkv), L.ptr(o1), L.ptr(dout), L.ptr(lse1), L.ptr(delta), L.ptr(dqkv), B, T, H, ctypes.c_float(0.1),
                              ctypes.c_uint64(11), L.stream_ptr()))
    direction = torch.randn_like(qkv.float())
    eps = 0.05
    op, _ = fwd((qkv.float() + eps * direction).bfloat16(), 0.1, 11)
    om, _ = fwd((qkv.float() - eps * direction).bfloat16(), 0.1, 11)
    fd = ((op.float() - om.float()) * dout.float()).sum().item() / (2 * eps)
    an = (dqkv.float() * direction).sum().item()
    assert abs(fd - an) / (abs(fd) + 1e-6) < 0.1


@pytest.mark.parametrize("d", [128, 512, 1024])
@pytest.mark.parametrize("dbl", [0, 1])
def test_layernorm_fwd_bwd(L, d, dbl):
    lib = L.lib()
    M = 333
    torch.manual_seed(2)
    x = torch.randn(M, d, device="cuda") * 2 + 0.3
    w1 = torch.randn(d, device="cuda") * 0.1 + 1; b1 = torch.randn(d, device="cuda") * 0.1
    w2 = torch.randn(d, device="cuda") * 0.1 + 1; b2 = torch.randn(d, device="cuda") * 0.1
    y = torch.zeros(M, d, device="cuda")
    stats = torch.zeros(M, 4 if dbl else 2, device="cuda")
    L.check(lib.ttts_layernorm_fwd(L.ptr(x), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(y), L.ptr(stats), M, d, dbl, 0, L.stream_ptr()))
    xr = x.clone().requires_grad_(True)
    ps = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    ref = torch.nn.functional.layer_norm(xr, (d,), ps[0], ps[1], 1e-5)
    if dbl:
        ref = torch.nn.functional.layer_norm(ref, (d,), ps[2], ps[3], 1e-5)
    assert rel(y, ref) < 1e-5
    dy = torch.randn(M, d, device="cuda")
    g_in = torch.randn(M, d, device="cuda")
    ref.backward(dy)
    g_out = torch.zeros(M, d, device="cuda"); g16 = torch.zeros(M, d, device="cuda", dtype=torch.bfloat16)
    grads = [torch.zeros(d, device="cuda") for _ in range(5)]
    L.check(lib.ttts_layernorm_bwd(L.ptr(dy), 1, L.ptr(x), L.ptr(stats), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(g_in), L.ptr(g_out), L.ptr(g16),
                                   L.ptr(grads[0]), L.ptr(grads[1]), L.ptr(grads[2]) if dbl else None, L.ptr(grads[3]) if dbl else None,
                                   L.ptr(grads[4]), M, d, dbl, L.stream_ptr()))
    assert rel(g_out, g_in + xr.grad) < 1e-4
    assert rel(g16, g_out) < 5e-3
    assert rel(grads[0], ps[0].grad) < 1e-4 and rel(grads[1], ps[1].grad) < 1e-4
    if dbl:
        assert rel(grads[2], ps[2].grad) < 1e-4 and rel(grads[3], ps[3].grad) < 1e-4
    assert rel(grads[4], g16.float().sum(0)) < 1e-4


@pytest.mark.parametrize("V,ld", [(1026, 1088), (257, 320)])
def test_cross_entropy_fwd_bwd(L, V, ld):
    lib = L.lib()
    rows = 777
    torch.manual_seed(4)
    logits = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)
    logits[:, :V] = (torch.randn(rows, V, device="cuda") * 2).bfloat16()
    logits[:, V:] = 50.0                                  # padding columns must be ignored
    tgt = torch.randint(0, V, (rows,), device="cuda", dtype=torch.int32)
    row_loss = torch.zeros(rows, device="cuda"); row_lse = torch.zeros(rows, device="cuda"); loss = torch.zeros(1, device="cuda")
    L.check(lib.ttts_ce_fwd(L.ptr(logits), ld, V, L.ptr(tgt), rows, L.ptr(row_loss), L.ptr(row_lse), L.ptr(loss), L.stream_ptr()))
    lr = logits[:, :V].float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, tgt.long())
    assert abs(loss.item() - ref.item()) < 1e-5
    gs = torch.tensor([0.37], device="cuda")
    (ref * 0.37 * 2.0).backward()
    dl = torch.full((rows, ld), 9.0, device="cuda", dtype=torch.bfloat16)
    L.check(lib.ttts_ce_bwd(L.ptr(logits), ld, V, L.ptr(tgt), rows, L.ptr(row_lse), L.ptr(gs), ctypes.c_float(2.0), L.ptr(dl), L.stream_ptr()))
    assert rel(dl[:, :V], lr.grad) < 5e-3
    assert bool((dl[:, V:] == 0).all())


@pytest.mark.parametrize("world,max_norm", [(1, 1.0), (8, 1.0), (1, 1e9)])
def test_adamw_kernel_exact(L, world, max_norm):
    """ttts_grad_norm + ttts_adamw_step vs clip_grad_norm_ + torch.optim.AdamW(betas (0.9, 0.96), wd 0.01) on IDENTICAL fp32 gradients
    (ttts/gpt/train.py:22-31,56,114-118), three steps: fp32 AdamW is exact arithmetic, so parameters, both moments and the norm agree to
    rounding (FMA contraction / lerp association: ~1 ulp per step), incl. grad_scale = 1/world on a SUMMED gradient buffer, the clip, the
    bias corrections, weight decay on every tensor, and the bf16 shadow == p.bfloat16()."""
    lib = L.lib()
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(7 + world)
    p = torch.randn(n, device="cuda", generator=g) * 0.05
    m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    norm = torch.zeros(1, device="cuda"); scratch = torch.zeros(2048, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=3e-4, betas=(0.9, 0.96), eps=1e-8, weight_decay=0.01)
    for step in range(1, 4):
        gsum = torch.randn(n, device="cuda", generator=g) * (0.01 * step) * world        # what the all-reduce (SUM) leaves in the buffer
        ref.grad = gsum / world
        want_norm = torch.nn.utils.clip_grad_norm_([ref], max_norm)
        opt.step()
        L.check(lib.ttts_grad_norm(gsum.data_ptr(), n, scratch.data_ptr(), norm.data_ptr(), L.stream_ptr().value))
        L.check(lib.ttts_adamw_step(p.data_ptr(), gsum.data_ptr(), m.data_ptr(), v.data_ptr(), p16.data_ptr(), n, norm.data_ptr(), max_norm,
                                    1.0 / world, 3e-4, 0.9, 0.96, 1e-8, 0.01, step, L.stream_ptr().value))
        assert abs(norm.item() / world - want_norm.item()) <= 2e-6 * want_norm.item()
        st = opt.state[ref]
        assert float((m - st["exp_avg"]).abs().max()) <= 2e-6 * float(st["exp_avg"].abs().max())
        assert float((v - st["exp_avg_sq"]).abs().max()) <= 2e-6 * float(st["exp_avg_sq"].abs().max())
        assert float((p - ref.data).abs().max()) <= 1e-6 * float(ref.data.abs().max())
        assert torch.equal(p16, p.bfloat16())
    assert rel(p, ref.data) < 1e-7
