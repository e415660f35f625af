"""CPU: the flow oracle against the REAL reference (tests/golden/flow.npz) and the training graph of the flow + KL term
(ttts_b200/vqvae/train_flow.py, next scope row) over the torch restatement of the kernel contract (tests/ref_kernels.py)."""
import os
import sys

import numpy as np
import torch

from oracle import disc_oracle as DO
from oracle import flow_oracle as FO
from ttts_b200.vqvae.train_encoder import Var
from ttts_b200.vqvae.train_flow import FlowGraph

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402


def _check_grads(z, grads):
    names = [str(n) for n in z["names"]]
    assert set(names) == set(grads.keys())
    floor = 1e-6 * float(np.sqrt((z["norm"] ** 2).sum()))
    for i, k in enumerate(names):
        gk = grads[k]
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i))
        scale = float(z["norm"][i])
        assert abs(float(gk.norm()) - scale) <= 2e-3 * scale + floor, (k, float(gk.norm()), scale)
        assert abs(float((gk * d).sum()) - float(z["proj"][i])) <= 1e-2 * scale + floor, k


def test_flow_oracle_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "flow.npz"))
    zz, ge, mask, logs_q, m_p, logs_p = FO.golden_inputs()
    assert abs(float(zz.sum()) - float(z["z_sum"])) < 1e-3                                  # the regenerated inputs are the minted ones
    P = {k: v.clone().requires_grad_(True) for k, v in FO.init_params(seed=6).items()}
    zz.requires_grad_(True); ge.requires_grad_(True)
    z_p = FO.flow(P, zz, mask, ge)
    assert np.abs(z_p.detach().numpy() - z["z_p"]).max() <= 2e-5 * np.abs(z["z_p"]).max()
    loss = DO.kl_loss(z_p, logs_q, m_p, logs_p, mask)
    assert abs(float(loss.detach()) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    loss.backward()
    _check_grads(z, {k: v.grad for k, v in P.items()})
    assert np.linalg.norm(zz.grad.numpy() - z["dz"]) <= 1e-4 * np.linalg.norm(z["dz"])
    assert np.linalg.norm(ge.grad.numpy() - z["dg"]) <= 1e-4 * np.linalg.norm(z["dg"])


def test_flow_training_graph_reproduces_the_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "flow.npz"))
    zz, ge, mask, logs_q, m_p, logs_p = FO.golden_inputs()
    graph = FlowGraph(TorchRefKernels(), FO.init_params(seed=6))
    zv, gv = Var(zz), Var(ge)
    mask2 = mask[:, 0].contiguous()
    z_p = graph.forward(zv, mask2, gv)
    assert np.abs(z_p.v.numpy() - z["z_p"]).max() <= 2e-5 * np.abs(z["z_p"]).max()
    lq, mp, lp = Var(logs_q), Var(m_p), Var(logs_p)
    loss = graph.ops.kl(z_p, lq, mp, lp, mask2)
    assert abs(float(loss.v) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    _check_grads(z, graph.backward(loss))
    assert np.linalg.norm(zv.g.numpy() - z["dz"]) <= 1e-4 * np.linalg.norm(z["dz"])
    assert np.linalg.norm(gv.g.numpy() - z["dg"]) <= 1e-4 * np.linalg.norm(z["dg"])
    # the prior-side inputs of the KL term receive their gradients too (they come from enc_p_2 / enc_q in the full step)
    a, b, c, d = [t.clone().requires_grad_(True) for t in (z_p.v, logs_q, m_p, logs_p)]
    DO.kl_loss(a, b, c, d, mask).backward()
    for var, ref in ((lq, b), (mp, c), (lp, d)):
        assert float((var.g - ref.grad).abs().max()) <= 1e-6 + 1e-5 * float(ref.grad.abs().max())
