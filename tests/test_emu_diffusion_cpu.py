"""CPU emulation of the diffusion kernels (tests/emu builds ttts_b200/csrc/diffusion_kernels.cu for the host) through the PRODUCT backend's
marshalling code (ttts_b200/diffusion/kernels.py) against the op contract tests/ref_kernels.py: GroupNorm (+ modulation, + SiLU) forward /
backward, SiLU, attention with the bucketed relative-position bias (head widths 8 / 16 / 32, lengths that are not tile multiples, several
query / key tiles) forward / backward, q_sample, and the loss incl. the t = 0 decoder-NLL branch and its |x| > 0.999 cases."""
import os
import shutil
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402

R = TorchRefKernels()


@pytest.fixture(scope="module")
def K(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import emu_kernels as EK
    return EK.emu_diffusion_kernels(EK.build_all(str(tmp_path_factory.mktemp("emu"))))


def close(got, want, tol=3e-5):
    assert got.shape == want.shape, (got.shape, want.shape)
    err = float((got - want).abs().max())
    assert err <= tol * max(1.0, float(want.abs().max())), (err, float(want.abs().max()))


@pytest.mark.parametrize("B,C,T,G,mod,silu", [(2, 32, 9, 8, False, False), (2, 64, 21, 16, True, True), (1, 128, 5, 32, False, True), (3, 16, 40, 8, True, False)])
def test_groupnorm_fwd_bwd(K, B, C, T, G, mod, silu):
    g = torch.Generator().manual_seed(B * 100 + C + T)
    x = torch.randn(B, C, T, generator=g) * 1.7 + 0.3
    gamma, beta = torch.rand(C, generator=g) + 0.5, 0.2 * torch.randn(C, generator=g)
    scale = 0.4 * torch.randn(B, C, 1, generator=g) if mod else None
    shift = 0.4 * torch.randn(B, C, 1, generator=g) if mod else None
    dy = torch.randn(B, C, T, generator=g)
    y, stats = K.gn_fwd(x, gamma, beta, G, scale, shift, silu)
    yr, sr = R.gn_fwd(x, gamma, beta, G, scale, shift, silu)
    close(y, yr); close(stats, sr)
    got = K.gn_bwd(dy, x, stats, gamma, beta, G, scale, shift, silu)
    want = R.gn_bwd(dy, x, sr, gamma, beta, G, scale, shift, silu)
    for a, b in zip(got, want):
        if b is None:
            assert a is None
        else:
            close(a, b, 5e-5)


def test_silu(K):
    x, dy = torch.randn(3, 7, 11) * 3, torch.randn(3, 7, 11)
    close(K.silu_fwd(x), R.silu_fwd(x))
    close(K.silu_bwd(dy, x), R.silu_bwd(dy, x))


@pytest.mark.parametrize("B,H,ch,T", [(1, 2, 16, 70), (2, 1, 32, 24), (1, 2, 8, 130), (1, 1, 64, 65)])
def test_attn_bias_fwd_bwd(K, B, H, ch, T):
    from ttts_b200.diffusion.train_graph import diagonal_buckets
    g = torch.Generator().manual_seed(7 * T + ch)
    qkv = torch.randn(B, 3 * H * ch, T, generator=g)
    table = 0.5 * torch.randn(32, H, generator=g)
    diag = diagonal_buckets(T)
    do = torch.randn(B, H * ch, T, generator=g)
    out, lse = K.attn_bias_fwd(qkv, table, H, diag)
    outr, lser = R.attn_bias_fwd(qkv, table, H, diag)
    close(out, outr); close(lse, lser)
    dqkv, dtab = K.attn_bias_bwd(do, qkv, out, lse, table, H, diag)
    dqr, dtr = R.attn_bias_bwd(do, qkv, outr, lser, table, H, diag)
    close(dqkv, dqr, 5e-5); close(dtab, dtr, 5e-5)


def test_q_sample_and_loss(K):
    from ttts_b200.diffusion.train_graph import coef_table
    g = torch.Generator().manual_seed(3)
    B, Cn, T = 4, 10, 37
    t = torch.tensor([0, 3, 500, 999])
    coef = coef_table(t)
    x0 = 0.6 * torch.randn(B, Cn, T, generator=g)
    x0[0, :2, :5] = -1.3; x0[0, 2:4, :5] = 1.2
    noise = torch.randn(B, Cn, T, generator=g)
    xt = K.q_sample(x0, noise, coef)
    close(xt, R.q_sample(x0, noise, coef), 1e-6)
    out = torch.randn(B, 2 * Cn, T, generator=g)
    out[0, Cn:] *= 2.0                                                 # variance values outside [-1, 1] too
    t0 = (t == 0).int()
    loss, (mse, vb) = K.diff_loss_fwd(out, x0, xt, noise, coef, t0)
    lr, (mr, vr) = R.diff_loss_fwd(out, x0, xt, noise, coef, t0)
    close(loss, lr, 1e-5); close(mse, mr, 1e-5); close(vb, vr, 2e-5)
    dL = torch.tensor([1.7])
    close(K.diff_loss_bwd(dL, out, x0, xt, noise, coef, t0), R.diff_loss_bwd(dL, out, x0, xt, noise, coef, t0), 5e-5)


@pytest.mark.parametrize("B,C,T", [(2, 40, 37), (1, 64, 64), (3, 7, 5), (2, 130, 70), (1, 65, 129)])
def test_layout_conversion_kernels(K, B, C, T):
    """ttts_cl_split / ttts_cl_unpack: [B,C,T] fp32 <-> position-major split-bf16 rows with zero rows between the clips"""
    g = torch.Generator().manual_seed(B + C + T)
    x = torch.randn(B, C, T, generator=g) * 3
    buf = K._cl_split("x", x)
    assert tuple(buf.shape) == (2 + B * (T + 1), 2 * C)
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    want = torch.zeros_like(buf)
    for b in range(B):
        want[1 + b * (T + 1):1 + b * (T + 1) + T, :C] = hi[b].t()
        want[1 + b * (T + 1):1 + b * (T + 1) + T, C:] = lo[b].t()
    assert torch.equal(buf, want)
    # hi + lo reproduces x to ~2^-16 relative
    rec = (buf[:, :C].float() + buf[:, C:].float())[1:1 + B * (T + 1)].reshape(B, T + 1, C)[:, :T].transpose(1, 2)
    assert float((rec - x).abs().max()) <= 2 ** -15 * float(x.abs().max())
    D = torch.randn(B * (T + 1), C + 8, generator=g)
    y = K._cl_unpack(D, B, C, T)
    want_y = D[:, :C].reshape(B, T + 1, C)[:, :T].transpose(1, 2).contiguous()
    assert torch.equal(y, want_y)
    # with the derivative of a leaky ReLU (slope 0.1) on the layer's input folded in, and the activation itself folded into the split
    yg = K._cl_unpack(D, B, C, T, lrelu_x=x)
    assert torch.equal(yg, torch.where(x > 0, want_y, 0.1 * want_y))
    xa = torch.nn.functional.leaky_relu(x, 0.1)
    buf2 = K._cl_split("x", x, lrelu=True).clone()
    assert torch.equal(buf2[:, :C], K._cl_split("x", xa)[:, :C])
    # other geometries: padding rows before each clip, a longer clip pitch
    buf3 = K._cl_split("g", x, T + 5, 3, 3 + B * (T + 5) + 2)
    want3 = torch.zeros_like(buf3)
    for b in range(B):
        want3[3 + b * (T + 5):3 + b * (T + 5) + T, :C] = hi[b].t()
        want3[3 + b * (T + 5):3 + b * (T + 5) + T, C:] = lo[b].t()
    assert torch.equal(buf3, want3)
    # de-interleaved order (dilation d): sample t of clip b -> sub-clip b d + t mod d, position t // d; and back
    for d in (3, 5):
        Ts = (T + d - 1) // d
        rpc = Ts + 2
        buf4 = K._cl_split("d%d" % d, x, rpc, 1, 1 + B * d * rpc + 1, False, d)
        want4 = torch.zeros_like(buf4)
        for b in range(B):
            for r in range(d):
                n = len(range(r, T, d))
                want4[1 + (b * d + r) * rpc:1 + (b * d + r) * rpc + n, :C] = hi[b, :, r::d].t()
                want4[1 + (b * d + r) * rpc:1 + (b * d + r) * rpc + n, C:] = lo[b, :, r::d].t()
        assert torch.equal(buf4, want4)
        D4 = torch.randn(1 + B * d * rpc + 1, C + 4, generator=g)
        y4 = K._cl_unpack(D4, B, C, T, rpc, 1, None, d)
        want_y4 = torch.empty(B, C, T)
        for b in range(B):
            for r in range(d):
                n = len(range(r, T, d))
                want_y4[b, :, r::d] = D4[1 + (b * d + r) * rpc:1 + (b * d + r) * rpc + n, :C].t()
        assert torch.equal(y4, want_y4)


@pytest.mark.parametrize("Cout,Cin,Kw", [(5, 7, 3), (64, 32, 11), (3, 130, 1)])
def test_weight_operands_of_the_tap_concatenated_route(K, Cout, Cin, Kw):
    """ttts_conv_w_concat: w [Cout,Cin,K] -> W1 = per tap [hi | hi], W2 = per tap [lo | 0]; the input-gradient form has the channel roles swapped
    and the taps reversed -- against the torch formulation the kernel replaced"""
    g = torch.Generator().manual_seed(Cout + Cin + Kw)
    w = torch.randn(Cout, Cin, Kw, generator=g)
    assert hasattr(K.lib, "ttts_conv_w_concat") and K.fused_wprep
    for flip_t in (False, True):
        W1, W2 = K._concat_weights(w, flip_t)
        K.fused_wprep = False
        try:
            R1, R2 = K._concat_weights(w, flip_t)
        finally:
            K.fused_wprep = True
        assert W1.shape == R1.shape and torch.equal(W1, R1) and torch.equal(W2, R2)
