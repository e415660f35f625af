"""CPU: the PRODUCT backend class (`CudaKernels`) driven through the host builds of its CUDA sources (tests/emu_kernels.py: EmuKernels) --
the same Python marshalling code, the same kernels, on the CPU emulation -- under the training graphs, against the REAL reference's goldens.
What this adds to the per-kernel emulation tests: the ctypes argument order / shapes / flags of every backend method, and the kernels under
the shapes the graphs actually produce."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import flow_oracle as FO  # noqa: E402
from oracle import text_encoder_oracle as TO  # noqa: E402
from ttts_b200.vqvae.train_encoder import Var  # noqa: E402


@pytest.fixture(scope="module")
def K(tmp_path_factory):
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    import emu_kernels
    return emu_kernels.EmuKernels(emu_kernels.build_all(str(tmp_path_factory.mktemp("emu"))))


def _check(z, grads, tol=3e-3):
    names = [str(n) for n in z["names"]]
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = grads[k]
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= tol * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 5 * tol * scale + floor, k


SLOW = pytest.mark.skipif(os.environ.get("TTTS_SLOW_EMU") != "1", reason="whole graphs at golden size on the emulation: >10 minutes of OS-thread churn; set TTTS_SLOW_EMU=1")


def _close(a, b, tol=1e-4):
    if a is None or b is None:
        assert a is None and b is None
        return
    assert tuple(a.shape) == tuple(b.shape), (a.shape, b.shape)
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


def test_every_backend_method_marshals_like_the_contract(K):
    """each method of the product backend, once, on tiny tensors: same results as the op contract (tests/ref_kernels.py).  Cheap, and it is the
    ctypes argument order / shapes / flags of EVERY call the GPU will see."""
    R = K.ref
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g)
    B, C, T = 2, 6, 9
    x, dy, w, b = rn(B, C, T), rn(B, 4, T), rn(4, C, 3) * 0.3, rn(4)
    for a_, b_ in zip(K.conv_bwd(dy, x, w, 1, 1, 1, True, True, True), R.conv_bwd(dy, x, w, 1, 1, 1, True, True, True)):
        _close(a_, b_)
    dys = rn(B, 4, 5)
    for a_, b_ in zip(K.conv_bwd(dys, x, w, 2, 1, 1, False, True, False), R.conv_bwd(dys, x, w, 2, 1, 1, False, True, False)):
        _close(a_, b_)
    wt, bt = rn(C, 4, 4) * 0.3, rn(4)                                      # ConvTranspose1d weight [Cin, Cout, K]
    _close(K.convT_fwd(x, wt, bt, 2, 1), R.convT_fwd(x, wt, bt, 2, 1))
    dyt = rn(*R.convT_fwd(x, wt, bt, 2, 1).shape)
    for a_, b_ in zip(K.convT_bwd(dyt, x, wt, 2, 1, True), R.convT_bwd(dyt, x, wt, 2, 1, True)):
        _close(a_, b_)
    xg, wg, bg = rn(B, 8, T), rn(6, 4, 3) * 0.3, rn(6)                     # grouped: 2 groups
    _close(K.conv_fwd(xg, wg, bg, 2, 1, 1, False, groups=2), R.conv_fwd(xg, wg, bg, 2, 1, 1, False, groups=2))
    dyg = rn(*R.conv_fwd(xg, wg, bg, 2, 1, 1, False, groups=2).shape)
    for a_, b_ in zip(K.conv_bwd(dyg, xg, wg, 2, 1, 1, False, True, True, groups=2), R.conv_bwd(dyg, xg, wg, 2, 1, 1, False, True, True, groups=2)):
        _close(a_, b_)
    v, gg, dw = rn(4, C, 3), torch.rand(4, 1, 1, generator=g) + 0.5, rn(4, C, 3)
    for a_, b_ in zip(K.wn_bwd(dw, v, gg), R.wn_bwd(dw, v, gg.reshape(-1))):
        _close(a_.reshape(-1), b_.reshape(-1))
    y2 = rn(B, C, T)
    mask = (torch.rand(B, T, generator=g) > 0.3).float()
    _close(K.add(x, y2), R.add(x, y2)); _close(K.scale(x, 0.3), R.scale(x, 0.3)); _close(K.mul_mask(x, mask), R.mul_mask(x, mask))
    raw, d3 = rn(B, 2 * C, T), rn(B, C, T)
    _close(K.glu_fwd(raw), R.glu_fwd(raw)); _close(K.glu_bwd(d3, raw), R.glu_bwd(d3, raw))
    _close(K.mish_fwd(x), R.mish_fwd(x)); _close(K.mish_bwd(d3, x), R.mish_bwd(d3, x))
    _close(K.lrelu_fwd(x, 0.1), R.lrelu_fwd(x, 0.1)); _close(K.lrelu_bwd(d3, x, 0.01), R.lrelu_bwd(d3, x, 0.01))
    _close(K.tanh_fwd(x), R.tanh_fwd(x)); _close(K.tanh_bwd(d3, x), R.tanh_bwd(d3, x))
    cond = rn(B, 2 * C)
    _close(K.gate_fwd(raw, cond), R.gate_fwd(raw, cond))
    for a_, b_ in zip(K.gate_bwd(d3, raw, cond), R.gate_bwd(d3, raw, cond)):
        _close(a_, b_)
    _close(K.gate_bwd(d3, raw, None)[0], R.gate_bwd(d3, raw, None)[0])
    cb = rn(B, C, 1)
    _close(K.add_bcast_fwd(x, cb), R.add_bcast_fwd(x, cb)); _close(K.add_bcast_bwd(d3), R.add_bcast_bwd(d3))
    from ttts_b200.vqvae.train_encoder import kaiser_sinc_filter12
    la, lb, filt = 0.3 * rn(C), 0.3 * rn(C), kaiser_sinc_filter12("cpu")
    for a_, b_ in zip(K.snake_bwd(d3, x, la, lb, filt), R.snake_bwd(d3, x, la, lb, filt)):
        _close(a_, b_)
    lens = torch.tensor([9, 4], dtype=torch.int64)
    q, k2, v2, do = rn(B, 8, T), rn(B, 8, T), rn(B, 8, T), rn(B, 8, T)
    for a_, b_ in zip(K.mha_bwd(do, q, k2, v2, lens, 2, 3.0), R.mha_bwd(do, q, k2, v2, lens, 2, 3.0)):
        _close(a_, b_)
    dmm = rn(B, C)
    _close(K.masked_mean_bwd(dmm, lens, T), R.masked_mean_bwd(dmm, lens, T))
    stats, eps, dz = rn(B, 2 * C, T), rn(B, C, T), rn(B, C, T)
    _close(K.posterior_bwd(dz, stats, eps, mask), R.posterior_bwd(dz, stats, eps, mask))
    dL = torch.tensor([0.7])
    _close(K.lsgan_fwd(x, 1.0), R.lsgan_fwd(x, 1.0)); _close(K.lsgan_bwd(dL, x, 1.0), R.lsgan_bwd(dL, x, 1.0))
    _close(K.l1_fwd(x, y2), R.l1_fwd(x, y2)); _close(K.l1_bwd(dL, x, y2), R.l1_bwd(dL, x, y2))
    zp, lq, mp, lp = rn(B, C, T), 0.3 * rn(B, C, T), rn(B, C, T), 0.3 * rn(B, C, T)
    _close(K.kl_fwd(zp, lq, mp, lp, mask), R.kl_fwd(zp, lq, mp, lp, mask))
    for a_, b_ in zip(K.kl_bwd(dL, zp, lq, mp, lp, mask), R.kl_bwd(dL, zp, lq, mp, lp, mask)):
        _close(a_, b_)
    ek, ev = rn(1, 5, 4) * 0.5, rn(1, 5, 4) * 0.5                         # window 2, dk 4, 2 heads
    _close(K.attn_fwd(q, k2, v2, ek, ev, lens, lens, 2), R.attn_fwd(q, k2, v2, ek, ev, lens, lens, 2))
    for a_, b_ in zip(K.attn_bwd(do, q, k2, v2, ek, ev, lens, lens, 2), R.attn_bwd(do, q, k2, v2, ek, ev, lens, lens, 2)):
        _close(a_, b_)
    kx, vx, klen = rn(B, 8, 5), rn(B, 8, 5), torch.tensor([5, 2], dtype=torch.int64)
    _close(K.attn_fwd(q, kx, vx, None, None, lens, klen, 2), R.attn_fwd(q, kx, vx, None, None, lens, klen, 2))
    for a_, b_ in zip(K.attn_bwd(do, q, kx, vx, None, None, lens, klen, 2)[:3], R.attn_bwd(do, q, kx, vx, None, None, lens, klen, 2)[:3]):
        _close(a_, b_)
    gam, bet = torch.rand(C, generator=g) + 0.5, rn(C)
    _close(K.lnc_fwd(x, gam, bet), R.lnc_fwd(x, gam, bet))
    for a_, b_ in zip(K.lnc_bwd(d3, x, gam, bet), R.lnc_bwd(d3, x, gam, bet)):
        _close(a_, b_)
    wav = torch.clamp(0.3 * rn(1, 2560), -1, 1)                            # 4 frames of the v2 front end
    dm = rn(1, 128, 4)
    got, want = K.logmel_bwd(dm, wav), R.logmel_bwd(dm, wav)
    assert float((got - want).norm()) <= 5e-4 * float(want.norm())


@SLOW
def test_flow_and_kl_through_the_product_backend(K, golden_dir):
    from ttts_b200.vqvae.train_flow import FlowGraph
    z = np.load(os.path.join(golden_dir, "flow.npz"))
    zz, ge, mask, logs_q, m_p, logs_p = FO.golden_inputs()
    graph = FlowGraph(K, FO.init_params(seed=6))
    zv, gv, mask2 = Var(zz), Var(ge), mask[:, 0].contiguous()
    z_p = graph.forward(zv, mask2, gv)
    assert np.abs(z_p.v.numpy() - z["z_p"]).max() <= 5e-5 * np.abs(z["z_p"]).max()
    loss = graph.ops.kl(z_p, Var(logs_q), Var(m_p), Var(logs_p), mask2)
    assert abs(float(loss.v) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    _check(z, graph.backward(loss))
    assert np.linalg.norm(zv.g.numpy() - z["dz"]) <= 1e-3 * np.linalg.norm(z["dz"])


@SLOW
def test_prior_encoder_through_the_product_backend(K, golden_dir):
    from ttts_b200.vqvae.train_text_encoder import TextEncoderGraph
    z = np.load(os.path.join(golden_dir, "text_encoder.npz"))
    y, y_lengths, text, text_lengths, ge = TO.golden_inputs()
    graph = TextEncoderGraph(K, TO.init_params(seed=8))
    yv, gv = Var(y), Var(ge)
    _, stats = graph.forward(yv, y_lengths, text, text_lengths, gv)
    assert np.abs(stats.v[:, :192].numpy() - z["m"]).max() <= 1e-4 * max(1.0, np.abs(z["m"]).max())
    gR = torch.Generator().manual_seed(62)
    R1, R2 = torch.randn(z["m"].shape, generator=gR), torch.randn(z["m"].shape, generator=gR)
    stats.g = torch.cat([R1, R2], dim=1)
    graph.tape.backward()
    _check(z, graph.grads())
    assert np.linalg.norm(yv.g.numpy() - z["dy"]) <= 1e-3 * np.linalg.norm(z["dy"])
