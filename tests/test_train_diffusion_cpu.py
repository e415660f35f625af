"""The diffusion training graph (ttts_b200/diffusion/train_graph.py) over the torch restatement of the kernel contract (tests/ref_kernels.py)
against the REAL reference's micro-step (tests/golden/diffusion.npz): pins the wiring of the graph -- model output, loss terms, and the gradient
of all 232 parameter tensors (incl. the exact zeros of the dropped layers)."""
import os

import numpy as np
import torch

from oracle import diffusion_oracle as DO
from ref_kernels import TorchRefKernels
from ttts_b200.diffusion.train_graph import DiffusionGraph, diagonal_buckets, nearest_index, coef_table


def check_against_golden(z, graph, lossv, terms, grads, tol=2e-4, out_tol=2e-5, to_cpu=lambda t: t):
    out = to_cpu(terms["model_out"].v).numpy()
    assert np.abs(out - z["model_out"]).max() <= out_tol * np.abs(z["model_out"]).max(), np.abs(out - z["model_out"]).max()
    assert np.allclose(to_cpu(terms["mse"]).numpy(), z["mse"], rtol=10 * out_tol, atol=0)
    # the t = 999 term is ~1e-6: a sum of O(1) quantities that cancel to fp32 rounding, so it carries an absolute floor
    assert np.allclose(to_cpu(terms["vb"]).numpy(), z["vb"], rtol=10 * out_tol, atol=1e-7)
    assert abs(float(lossv.v) - float(z["loss"])) <= 10 * out_tol * abs(float(z["loss"]))
    names = [str(n) for n in z["names"]]
    assert set(names) == set(grads.keys())
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = to_cpu(grads[k])
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        if scale == 0.0:
            assert float(gk.abs().max()) == 0.0, k
            continue
        assert abs(float(gk.norm()) - scale) <= tol * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= tol * scale + floor, k


def test_graph_wiring_vs_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "diffusion.npz"))
    cfg = DO.default_config(**DO.GOLDEN_CFG)
    graph = DiffusionGraph(TorchRefKernels(), DO.init_params(cfg, seed=12), cfg)
    I = DO.golden_inputs()
    lossv, terms = graph.loss(I["x_start"], torch.tensor(I["t"]), I["noise"], I["latent"], I["refer"], I["uncond"], I["dropped"])
    grads = graph.backward(lossv)
    check_against_golden(z, graph, lossv, terms, grads)


def test_host_tables_match_the_oracle():
    for T in (1, 7, 24, 200, 1024):
        d = diagonal_buckets(T)
        b = DO.rel_pos_bucket(T, T)
        idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
        assert torch.equal(d.long()[idx], b), T
    assert torch.equal(nearest_index(7, 24), DO.nearest_index(7, 24))
    t = [0, 1, 517, 999]
    assert torch.equal(coef_table(torch.tensor(t)), DO.coef_table(t))


def reference_loop(P0, cfg, batches, lr, steps):
    """the loop body of ttts/diffusion/train.py:156-196 with torch's own autograd / clip / AdamW / LambdaLR over the pinned oracle"""
    from ttts_b200.diffusion.train_step import warmup
    P = {k: v.clone().requires_grad_(True) for k, v in P0.items()}
    opt = torch.optim.AdamW(list(P.values()), lr=lr, betas=(0.9, 0.999), weight_decay=0.01)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=warmup)
    losses, norms = [], []
    for s in range(steps):
        b = batches[s]
        loss, _ = DO.training_loss(P, cfg, b["x_start"], b["t"], b["noise"], b["latent"], b["refer"], b["uncond"], b["dropped"])
        loss.backward()
        for p in P.values():                                        # `extraneous_addition * 0`: every parameter has a (zero) gradient
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        norms.append(float(torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)))
        opt.step(); opt.zero_grad(); sched.step()
        losses.append(float(loss))
    return losses, norms, {k: v.detach() for k, v in P.items()}


def step_batches(n, seed=5):
    out = []
    for s in range(n):
        I = DO.golden_inputs(seed=seed + s)
        I["t"] = [3 + s, 400 + 7 * s, 990 - s]
        I["dropped"] = (1, 4) if s % 2 == 0 else (2,)
        out.append(I)
    return out


def test_step_order_clip_adamw_warmup_vs_torch():
    """DiffusionStep (graph over the op contract + clip + AdamW + warm-up) against torch's own loop over the pinned oracle, 3 steps with a
    learning rate large enough for the warm-up (0, lr/1000, 2 lr/1000) to move the loss"""
    from ref_kernels import TorchAdamWClip
    from ttts_b200.diffusion.train_step import DiffusionStep
    cfg = DO.default_config(**DO.GOLDEN_CFG)
    P0 = DO.init_params(cfg, seed=12)
    batches = step_batches(3)
    lr = 2.0
    want_l, want_n, want_P = reference_loop(P0, cfg, batches, lr, 3)
    ds = DiffusionStep(TorchRefKernels(), P0, cfg, lr=lr, optimizer=TorchAdamWClip)
    for s in range(3):
        b = dict(batches[s]); b["t"] = torch.tensor(b["t"])
        out = ds.step([b])
        assert abs(float(out["loss"]) - want_l[s]) <= 1e-4 * abs(want_l[s]), (s, float(out["loss"]), want_l[s])
        assert abs(float(out["grad_norm"]) - want_n[s]) <= 1e-3 * want_n[s], (s, float(out["grad_norm"]), want_n[s])
    assert want_l[2] != want_l[0]
    got = ds.opt.params()
    num = sum(float((got[k] - want_P[k]).norm() ** 2) for k in want_P)
    den = sum(float((want_P[k] - P0[k]).norm() ** 2) for k in want_P)
    assert den > 0 and (num / den) ** 0.5 <= 2e-3, (num / den) ** 0.5
