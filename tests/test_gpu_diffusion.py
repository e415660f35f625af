"""GPU (-m gpu): the diffusion mel-refiner train step (BASELINE config 5) through the C ABI -- the new kernels op by op
against the contract (tests/ref_kernels.py, itself the autograd of the pinned restatements), the training graph against the REAL reference's
micro-step (tests/golden/diffusion.npz), the optimisation step against torch's own loop over the pinned oracle, and size-independent
properties at the BASELINE size (batch 32 x 1024 frames, 512 channels, 16 heads)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import diffusion_oracle as DO
from ref_kernels import TorchRefKernels

R = TorchRefKernels()


@pytest.fixture(scope="module")
def K():
    from ttts_b200.diffusion.kernels import DiffusionCudaKernels
    return DiffusionCudaKernels()


def close(got, want, tol=3e-5):
    got = got.cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    err = float((got - want).abs().max())
    assert err <= tol * max(1.0, float(want.abs().max())), (err, float(want.abs().max()))


def cu(*ts):
    return [t.cuda() if t is not None else None for t in ts]


@pytest.mark.parametrize("B,C,T,G,mod,silu", [(2, 32, 9, 8, False, False), (2, 64, 21, 16, True, True), (3, 128, 50, 32, False, True),
                                              (2, 512, 1024, 32, True, True), (2, 512, 257, 32, False, False)])
def test_groupnorm_fwd_bwd(K, B, C, T, G, mod, silu):
    g = torch.Generator().manual_seed(B * 100 + C + T)
    x = torch.randn(B, C, T, generator=g) * 1.7 + 0.3
    gamma, beta = torch.rand(C, generator=g) + 0.5, 0.2 * torch.randn(C, generator=g)
    scale = 0.4 * torch.randn(B, C, 1, generator=g) if mod else None
    shift = 0.4 * torch.randn(B, C, 1, generator=g) if mod else None
    dy = torch.randn(B, C, T, generator=g)
    xc, gc, bc, sc, hc, dyc = cu(x, gamma, beta, scale, shift, dy)
    y, stats = K.gn_fwd(xc, gc, bc, G, sc, hc, silu)
    yr, sr = R.gn_fwd(x, gamma, beta, G, scale, shift, silu)
    close(y, yr); close(stats, sr)
    got = K.gn_bwd(dyc, xc, stats, gc, bc, G, sc, hc, silu)
    want = R.gn_bwd(dy, x, sr, gamma, beta, G, scale, shift, silu)
    for a, b in zip(got, want):
        if b is None:
            assert a is None
        else:
            close(a, b, 2e-4 if T >= 257 else 5e-5)        # long reductions: fp32 summation order


def test_silu(K):
    x, dy = torch.randn(3, 70, 111) * 3, torch.randn(3, 70, 111)
    close(K.silu_fwd(x.cuda()), R.silu_fwd(x))
    close(K.silu_bwd(dy.cuda(), x.cuda()), R.silu_bwd(dy, x))


@pytest.mark.parametrize("B,H,ch,T", [(1, 2, 16, 70), (2, 1, 32, 24), (1, 2, 8, 130), (1, 1, 64, 65), (2, 3, 32, 256), (1, 2, 64, 232), (1, 2, 32, 1024)])
def test_attn_bias_fwd_bwd(K, B, H, ch, T):
    from ttts_b200.diffusion.train_graph import diagonal_buckets
    g = torch.Generator().manual_seed(7 * T + ch)
    qkv = torch.randn(B, 3 * H * ch, T, generator=g)
    table = 0.5 * torch.randn(32, H, generator=g)
    diag = diagonal_buckets(T)
    do = torch.randn(B, H * ch, T, generator=g)
    out, lse = K.attn_bias_fwd(qkv.cuda(), table.cuda(), H, diag.cuda())
    outr, lser = R.attn_bias_fwd(qkv, table, H, diag)
    close(out, outr); close(lse, lser)
    dqkv, dtab = K.attn_bias_bwd(do.cuda(), qkv.cuda(), out, lse, table.cuda(), H, diag.cuda())
    dqr, dtr = R.attn_bias_bwd(do, qkv, outr, lser, table, H, diag)
    close(dqkv, dqr, 1e-4); close(dtab, dtr, 2e-4)
    # deterministic: a second run gives the same bits
    dq2, dt2 = K.attn_bias_bwd(do.cuda(), qkv.cuda(), out, lse, table.cuda(), H, diag.cuda())
    assert torch.equal(dq2, dqkv) and torch.equal(dt2, dtab)


def test_q_sample_and_loss(K):
    from ttts_b200.diffusion.train_graph import coef_table
    g = torch.Generator().manual_seed(3)
    B, Cn, T = 4, 100, 137
    t = torch.tensor([0, 3, 500, 999])
    coef = coef_table(t)
    x0 = 0.6 * torch.randn(B, Cn, T, generator=g)
    x0[0, :20, :50] = -1.3; x0[0, 20:40, :50] = 1.2
    noise = torch.randn(B, Cn, T, generator=g)
    xt = K.q_sample(x0.cuda(), noise.cuda(), coef.cuda())
    xtr = R.q_sample(x0, noise, coef)
    close(xt, xtr, 1e-6)
    out = torch.randn(B, 2 * Cn, T, generator=g)
    out[0, Cn:] *= 2.0
    t0 = (t == 0).int()
    loss, (mse, vb) = K.diff_loss_fwd(out.cuda(), x0.cuda(), xt, noise.cuda(), coef.cuda(), t0.cuda())
    lr, (mr, vr) = R.diff_loss_fwd(out, x0, xtr, noise, coef, t0)
    close(loss, lr, 2e-5); close(mse, mr, 2e-5); close(vb, vr, 5e-5)
    dL = torch.tensor([1.7])
    close(K.diff_loss_bwd(dL.cuda(), out.cuda(), x0.cuda(), xt, noise.cuda(), coef.cuda(), t0.cuda()), R.diff_loss_bwd(dL, out, x0, xtr, noise, coef, t0), 1e-4)


def test_training_graph_vs_reference_golden(K, golden_dir):
    """DiffusionGraph over the CUDA kernels against the REAL reference's micro-step: model output, mse / vb terms, loss, and the gradient of
    all 232 parameter tensors (exact zeros for the dropped layers)"""
    from test_train_diffusion_cpu import check_against_golden
    from ttts_b200.diffusion.train_graph import DiffusionGraph
    z = np.load(os.path.join(golden_dir, "diffusion.npz"))
    cfg = DO.default_config(**DO.GOLDEN_CFG)
    graph = DiffusionGraph(K, {k: v.cuda() for k, v in DO.init_params(cfg, seed=12).items()}, cfg)
    I = DO.golden_inputs()
    lossv, terms = graph.loss(I["x_start"].cuda(), torch.tensor(I["t"]), I["noise"].cuda(), I["latent"].cuda(), I["refer"].cuda(), I["uncond"], I["dropped"])
    grads = graph.backward(lossv)
    check_against_golden(z, graph, lossv, terms, grads, tol=5e-4, out_tol=3e-5, to_cpu=lambda t: t.detach().cpu())


def test_optimisation_steps_vs_torch_loop(K):
    """DiffusionStep with the fused clip + AdamW kernels against torch's own loop (autograd of the pinned oracle, clip_grad_norm_, AdamW,
    LambdaLR warm-up) over 3 steps"""
    from test_train_diffusion_cpu import reference_loop, step_batches
    from ttts_b200.diffusion.train_step import DiffusionStep
    cfg = DO.default_config(**DO.GOLDEN_CFG)
    P0 = DO.init_params(cfg, seed=12)
    batches = step_batches(3)
    lr = 2.0
    want_l, want_n, want_P = reference_loop(P0, cfg, batches, lr, 3)
    ds = DiffusionStep(K, {k: v.cuda() for k, v in P0.items()}, cfg, lr=lr)
    for s in range(3):
        b = {k: (v.cuda() if torch.is_tensor(v) and k != "uncond" else v) for k, v in batches[s].items()}
        b["t"] = torch.tensor(b["t"])
        out = ds.step([b])
        assert abs(float(out["loss"]) - want_l[s]) <= 2e-4 * abs(want_l[s]), (s, float(out["loss"]), want_l[s])
        assert abs(float(out["grad_norm"]) - want_n[s]) <= 2e-3 * want_n[s], (s, float(out["grad_norm"]), want_n[s])
    got = ds.opt.params()
    num = sum(float((got[k].cpu() - want_P[k]).norm() ** 2) for k in want_P)
    den = sum(float((want_P[k] - P0[k]).norm() ** 2) for k in want_P)
    assert den > 0 and (num / den) ** 0.5 <= 5e-3, (num / den) ** 0.5


def test_full_size_properties(K):
    """BASELINE config 5 size (batch 4 of the 32 x 1024-frame shape, 512 channels, 16 heads, 6 layers): the loss is finite, per-sample terms do
    not depend on the other samples of the batch (no cross-batch leakage in GroupNorm / attention), the graph is run-to-run reproducible,
    and a finite-difference step along the gradient matches <g, g>"""
    from ttts_b200.diffusion.train_graph import DiffusionGraph
    cfg = DO.default_config()
    P = {k: v.cuda() for k, v in DO.init_params(cfg, seed=3).items()}
    g = torch.Generator().manual_seed(11)
    B, T, TL, TR = 4, 1024, 256, 200
    x0 = (0.5 * torch.randn(B, 100, T, generator=g)).cuda()
    noise = torch.randn(B, 100, T, generator=g).cuda()
    latent = torch.randn(B, 512, TL, generator=g).cuda()
    refer = (0.5 * torch.randn(B, 100, TR, generator=g)).cuda()
    t = torch.tensor([5, 250, 700, 999])

    def run(sel, params=P):
        graph = DiffusionGraph(K, params, cfg)
        lossv, terms = graph.loss(x0[sel].contiguous(), t[sel], noise[sel].contiguous(), latent[sel].contiguous(), refer[sel].contiguous(), None, (2,))
        return graph, lossv, terms
    graph, lossv, terms = run(slice(0, 4))
    assert torch.isfinite(lossv.v).all()
    per = (terms["mse"] + terms["vb"]).cpu()
    _, _, terms2 = run(slice(1, 3))
    per2 = (terms2["mse"] + terms2["vb"]).cpu()
    assert torch.allclose(per[1:3], per2, rtol=1e-5, atol=1e-7), (per, per2)
    grads = graph.backward(lossv)
    graph_b, lossv_b, _ = run(slice(0, 4))
    grads_b = graph_b.backward(lossv_b)
    assert torch.equal(lossv.v, lossv_b.v)                          # the forward is bit-reproducible
    # the weight-gradient kernel combines its position slices with fp32 atomics: run-to-run agreement to rounding, not to the bit
    for k in grads:
        assert float((grads[k] - grads_b[k]).abs().max()) <= 1e-5 * max(float(grads[k].abs().max()), 1e-12), k
    assert sum(float((v.double() ** 2).sum()) for v in grads.values()) > 0


def test_full_width_vs_oracle(K):
    """the BASELINE model (512 channels, 16 heads of 32, 6 + 3 layers, RefEncoder heads of 64) at a size the fp32 CPU oracle finishes in
    seconds (batch 2 x 256 frames, latent 64, reference 50): loss terms, model output and every gradient tensor.  (A finite-difference
    check of the loss is NOT a valid test here: the variational-bound term detaches the mean prediction, utils/diffusion.py:980.)"""
    from ttts_b200.diffusion.train_graph import DiffusionGraph
    cfg = DO.default_config()
    P0 = DO.init_params(cfg, seed=3)
    g = torch.Generator().manual_seed(21)
    B, T, TL, TR = 2, 256, 64, 50
    x0 = 0.5 * torch.randn(B, 100, T, generator=g)
    noise = torch.randn(B, 100, T, generator=g)
    latent = torch.randn(B, 512, TL, generator=g)
    refer = 0.5 * torch.randn(B, 100, TR, generator=g)
    t = [40, 800]
    uncond, dropped = torch.tensor([False, True]), (3, 7)
    P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
    x_t = DO.q_sample(x0, t, noise)
    out_ref = DO.model_forward(P, cfg, x_t, torch.tensor(t), latent, refer, uncond, dropped)
    mse_ref, vb_ref = DO.loss_terms(out_ref, x0, x_t, noise, t)
    loss_ref = (mse_ref + vb_ref).mean()
    loss_ref.backward()
    graph = DiffusionGraph(K, {k: v.cuda() for k, v in P0.items()}, cfg)
    lossv, terms = graph.loss(x0.cuda(), torch.tensor(t), noise.cuda(), latent.cuda(), refer.cuda(), uncond, dropped)
    grads = graph.backward(lossv)
    o = terms["model_out"].v.cpu()
    assert float((o - out_ref.detach()).abs().max()) <= 1e-4 * float(out_ref.detach().abs().max())
    assert torch.allclose(terms["mse"].cpu(), mse_ref.detach(), rtol=1e-4) and torch.allclose(terms["vb"].cpu(), vb_ref.detach(), rtol=1e-3, atol=1e-7)
    assert abs(float(lossv.v) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    num = den = 0.0
    for k, p in P.items():
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        gk = grads[k].cpu()
        n_ref = float(gr.norm())
        if n_ref == 0.0:
            assert float(gk.abs().max()) == 0.0, k
            continue
        e = float((gk - gr).norm())
        assert e <= 2e-3 * n_ref + 1e-7, (k, e, n_ref)
        num += e * e; den += n_ref * n_ref
    assert (num / den) ** 0.5 <= 5e-4, (num / den) ** 0.5


def test_trainer_surface_save_load(K, tmp_path):
    """ttts_b200.diffusion.train.Trainer (the surface of ttts/diffusion/train.py:78-256) on batches in DiffusionCollater's layout: the warm-up
    makes step 0 a no-op (lr 0), later steps move the parameters, a checkpoint {'step', 'model'} round-trips with the reference's keys"""
    from ttts_b200.diffusion.params import param_shapes
    from ttts_b200.diffusion.train import Trainer
    cfg = DO.default_config(**DO.GOLDEN_CFG)

    class FrozenGPT:                                               # stands in for UnifiedVoice(return_latent=True): [B, CL, in_latent_channels]
        mel_length_compression = 1024

        def __call__(self, text, tl, codes, wl, return_latent=True, clip_inputs=False):
            g = torch.Generator().manual_seed(int(codes.sum()) % 1000)
            return torch.randn(codes.shape[0], codes.shape[1], cfg["in_latent_channels"], generator=g).cuda()
    g = torch.Generator().manual_seed(0)
    batches = [dict(padded_text=torch.randint(0, 200, (3, 9), generator=g), padded_mel_code=torch.randint(0, 1024, (3, 6), generator=g),
                    padded_mel=2 * torch.randn(3, 100, 24, generator=g) - 5, padded_mel_refer=2 * torch.randn(3, 100, 10, generator=g) - 5) for _ in range(2)]
    batches.insert(1, None)                                        # a batch whose samples all failed to load is skipped (train.py:158-159)
    P0 = {k: v.cuda() for k, v in DO.init_params(cfg, seed=12).items()}
    tr = Trainer(FrozenGPT(), batches, cfg=cfg, lr=1.0, train_steps=3, params=P0)
    logs = []
    tr.train(log=lambda s, out: logs.append((s, float(out["loss"]), float(out["grad_norm"]))))
    assert [s for s, _, _ in logs] == [1, 2, 3] and all(np.isfinite(l) and np.isfinite(n) for _, l, n in logs)
    sd = tr.state_dict()
    assert set(sd.keys()) == set(param_shapes(cfg).keys()) and all(tuple(sd[k].shape) == param_shapes(cfg)[k] for k in sd)
    assert any(float((sd[k] - P0[k]).abs().max()) > 0 for k in sd)
    path = str(tmp_path / "model-0.pt")
    tr.save(path)
    tr2 = Trainer(FrozenGPT(), batches, cfg=cfg, lr=1.0, train_steps=3, params={k: torch.zeros_like(v) for k, v in P0.items()})
    tr2.load(path)
    assert tr2.step == 3 and all(torch.equal(tr2.state_dict()[k], sd[k]) for k in sd)


@pytest.mark.parametrize("B,Cin,T,Cout,Kw,stride,dil,pad,lrelu", [
    (4, 512, 600, 512, 3, 1, 1, 1, False), (2, 1024, 1030, 512, 1, 1, 1, 0, False), (3, 128, 700, 192, 3, 1, 1, 1, False), (2, 512, 1024, 1536, 1, 1, 1, 0, False),
    (10, 512, 760, 1024, 5, 3, 1, 2, False),      # DiscriminatorP 512 -> 1024, kernel 5, stride 3 (vq2.py:418-460)
    (16, 1024, 253, 1024, 5, 1, 1, 2, False),     # ... 1024 -> 1024, stride 1
    (4, 128, 2560, 128, 11, 1, 5, 25, True),      # Generator ResBlock1: kernel 11, dilation 5, leaky ReLU on the input (modules.py:224-318)
    (6, 256, 700, 128, 7, 2, 3, 9, True),         # stride, dilation and padding together
    (5, 192, 1000, 512, 7, 1, 1, 3, False),       # Generator.conv_pre
    (3, 64, 1700, 64, 11, 1, 1, 5, True),         # narrow layers (tap-concatenated form only): Generator ResBlocks at 64 ...
    (3, 32, 2100, 32, 7, 1, 1, 3, True),          # ... and 32 channels
    (4, 32, 1500, 128, 5, 3, 1, 2, False),        # DiscriminatorP 32 -> 128, stride 3
    (2, 64, 1203, 32, 3, 1, 1, 1, False),
    (3, 64, 1700, 64, 7, 1, 3, 9, True),          # dilated narrow layers: de-interleaved sub-clips (1700 is not a multiple of 3)
    (3, 32, 2100, 32, 3, 1, 5, 5, True),
    (2, 128, 1001, 256, 11, 1, 3, 15, True),      # dilated wide layer, de-interleaved
    (2, 128, 1000, 128, 3, 1, 3, 2, False)])      # padding not a multiple of the dilation: one GEMM pair per tap
def test_tensor_core_convolution_vs_torch(K, monkeypatch, B, Cin, T, Cout, Kw, stride, dil, pad, lrelu):
    """the split-bf16 tcgen05 GEMM route of the wide convolutions (forward, input gradient, weight and bias gradient; any kernel size, stride,
    dilation, padding, optional leaky ReLU on the input) against torch's fp64 convolution: fp32-grade agreement (three bf16 products per fp32
    product), and the route is actually taken for these shapes"""
    F = torch.nn.functional
    g = torch.Generator().manual_seed(Cin + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, Kw, generator=g) / (Cin * Kw) ** 0.5
    b = torch.randn(Cout, generator=g)
    xc, wc, bc = cu(x, w, b)
    if Cin % 64 or Cout % 64 or min(Cin, Cout) < 128:
        if not (K.tap_concat and K.wgrad_concat):
            pytest.skip("narrow layers take the route in the tap-concatenated form only")
        monkeypatch.setattr(K, "gemm_narrow", True)                   # opt-in in the product (TTTS_GEMM_NARROW=1)
    monkeypatch.setattr(K, "GEMM_MIN_POSITIONS", 1)                   # the product takes this route from 4096 output positions; the test shapes are smaller
    assert K._gemm_ok(xc, wc, stride, dil, pad, 1)
    xr, wr, br = x.cuda().double().requires_grad_(True), w.cuda().double().requires_grad_(True), b.cuda().double().requires_grad_(True)
    yr = F.conv1d(F.leaky_relu(xr, 0.1) if lrelu else xr, wr, br, stride=stride, dilation=dil, padding=pad)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy.cuda().double())
    y = K.conv_fwd(xc, wc, bc, stride, dil, pad, lrelu)
    dx, dw, db = K.conv_bwd(dy.cuda(), xc, wc, stride, dil, pad, lrelu, True, True)
    rel = lambda a, r: float((a.double() - r.double()).norm() / r.double().norm())
    assert y.shape == yr.shape
    assert rel(y, yr) <= 2e-5, rel(y, yr)
    assert rel(dx, xr.grad) <= 2e-5, rel(dx, xr.grad)
    assert rel(dw, wr.grad) <= 2e-5, rel(dw, wr.grad)
    assert rel(db, br.grad) <= 1e-5


@pytest.mark.parametrize("B,Cin,T,Cout,Kw,stride,pad", [(8, 512, 40, 256, 16, 8, 4), (4, 128, 700, 64, 4, 2, 1), (3, 256, 330, 128, 16, 8, 4)])
def test_conv_transpose_on_the_gemm_route(K, monkeypatch, B, Cin, T, Cout, Kw, stride, pad):
    """Generator.ups (ConvTranspose1d, modules.py / vq2.py Generator): forward = the phases of a strided input gradient, each a stride-1
    convolution with K / stride taps on the split-bf16 GEMM route; backward = a strided convolution (dx) and its weight gradient with the roles
    of x and dy swapped -- against torch's fp64 conv_transpose1d"""
    F = torch.nn.functional
    g = torch.Generator().manual_seed(Cin + T)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cin, Cout, Kw, generator=g) / (Cin * Kw / stride) ** 0.5
    b = torch.randn(Cout, generator=g)
    monkeypatch.setattr(K, "GEMM_MIN_POSITIONS", 1)
    monkeypatch.setattr(K, "gemm_narrow", True)
    xc, wc, bc = cu(x, w, b)
    xr, wr, br = x.cuda().double().requires_grad_(True), w.cuda().double().requires_grad_(True), b.cuda().double().requires_grad_(True)
    yr = F.conv_transpose1d(xr, wr, br, stride=stride, padding=pad)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy.cuda().double())
    y = K.convT_fwd(xc, wc, bc, stride, pad)
    dx, dw, db = K.convT_bwd(dy.cuda(), xc, wc, stride, pad, True)
    rel = lambda a, r: float((a.double() - r.double()).norm() / r.double().norm())
    assert y.shape == yr.shape and dx.shape == x.shape and dw.shape == w.shape
    assert rel(y, yr) <= 2e-5, rel(y, yr)
    assert rel(dx, xr.grad) <= 2e-5, rel(dx, xr.grad)
    assert rel(dw, wr.grad) <= 2e-5, rel(dw, wr.grad)
    assert rel(db, br.grad) <= 1e-5
