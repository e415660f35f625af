"""CPU: the training graphs COMPOSED on one tape (next scope row): latent z -> flow -> KL, and z -> Generator -> MultiPeriodDiscriminator ->
generator_loss + feature_loss, with z and the style vector g shared between the branches -- the dependency structure of the generator step
(ttts/vqvae/vq2.py:855-862, train.py:386-395) minus the parts not built yet.  Over the torch restatement of the kernel contract, against
torch.autograd through the pinned oracles (decoder / disc / flow oracles each equal the REAL reference)."""
import os
import sys

import torch

from oracle import decoder_oracle as DEC
from oracle import disc_oracle as DIS
from oracle import flow_oracle as FO
from ttts_b200.vqvae.train_decoder import DecoderGraph
from ttts_b200.vqvae.train_disc import DiscriminatorGraph
from ttts_b200.vqvae.train_encoder import Ops, Tape, Var
from ttts_b200.vqvae.train_flow import FlowGraph

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_kernels import TorchRefKernels  # noqa: E402


def test_graphs_compose_on_one_tape():
    K = TorchRefKernels()
    g0 = torch.Generator().manual_seed(3)
    B, T = 2, 6
    z0, ge0 = torch.randn(B, 192, T, generator=g0), torch.randn(B, 512, 1, generator=g0)
    mask = torch.ones(B, 1, T)
    logs_q, m_p, logs_p = [0.3 * torch.randn(B, 192, T, generator=g0) for _ in range(3)]
    y_real = torch.tanh(torch.randn(B, 1, T * 640, generator=g0))
    Pf, Pd, Pm = FO.init_params(seed=6), DEC.init_params(seed=9), DIS.init_params(seed=4)
    params = {**{"flow." + k: v for k, v in Pf.items()}, **{"dec." + k: v for k, v in Pd.items()}, **{"net_d." + k: v for k, v in Pm.items()}}

    # ---- ours: one tape, three graphs, shared leaves z and g ----
    tape = Tape()
    flow, dec, disc = FlowGraph(K, params, tape, "flow."), DecoderGraph(K, params, tape, "dec."), DiscriminatorGraph(K, params, tape, "net_d.")
    ops = Ops(K, tape)
    z, ge = Var(z0), Var(ge0)
    mask2 = mask[:, 0].contiguous()
    loss_kl = ops.kl(flow.forward(z, mask2, ge), Var(logs_q), Var(m_p), Var(logs_p), mask2)
    y_hat = dec.forward(z, ge)
    _, fmap_r = disc.forward(y_real)
    gen, fmap_g = disc.forward(y_hat)
    loss_gen, loss_fm = disc.generator_losses(gen, fmap_r, fmap_g)
    total = ops.add(ops.add(loss_gen, loss_fm), loss_kl)
    total.g = torch.ones(1)
    tape.backward()

    # ---- torch.autograd through the oracles ----
    zt, gt = z0.clone().requires_grad_(True), ge0.clone().requires_grad_(True)
    Pf_t = {k: v.clone().requires_grad_(True) for k, v in Pf.items()}
    Pd_t = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
    want_kl = DIS.kl_loss(FO.flow(Pf_t, zt, mask, gt), logs_q, m_p, logs_p, mask)
    yh = DEC.generator(Pd_t, zt, gt)
    _, y_d_g, fr, fg = DIS.mpd(Pm, y_real, yh)
    want = DIS.generator_loss(y_d_g) + DIS.feature_loss(fr, fg) + want_kl
    want.backward()
    assert abs(float(total.v) - float(want.detach())) <= 1e-5 * abs(float(want.detach()))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    assert rel(z.g, zt.grad) <= 1e-4 and rel(ge.g, gt.grad) <= 1e-4                        # both branches accumulate into the shared leaves
    for k, v in Pf_t.items():
        assert rel(flow.P[k].g, v.grad) <= 1e-3 + 1e-6 / (float(v.grad.norm()) + 1e-30), k
    for k, v in Pd_t.items():
        assert rel(dec.P[k].g.reshape(v.shape), v.grad) <= 1e-3, k


def test_mel_reconstruction_loss_through_the_generator():
    """loss_mel = l1(y_mel, mel_spectrogram_torch(y_hat)) * c_mel (train.py:357-366,389) with y_hat = Generator(z): tape vs torch.autograd"""
    K = TorchRefKernels()
    g0 = torch.Generator().manual_seed(8)
    B, T = 2, 5
    z0 = torch.randn(B, 192, T, generator=g0)
    y_real = torch.tanh(torch.randn(B, T * 640, generator=g0))
    Pd = DEC.init_params(seed=9)
    tape = Tape()
    dec, ops = DecoderGraph(K, Pd, tape), Ops(K, tape)
    z = Var(z0)
    y_hat = dec.forward(z, None)
    mel_hat = ops.logmel(ops.reshape(y_hat, (B, T * 640)))
    y_mel = K.logmel_fwd(y_real)
    loss = ops.scale(ops.l1_mean(y_mel, mel_hat), 45.0)
    loss.g = torch.ones(1)
    tape.backward()
    zt = z0.clone().requires_grad_(True)
    Pt = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
    want = torch.nn.functional.l1_loss(y_mel, K._logmel(DEC.generator(Pt, zt, None).squeeze(1))) * 45.0
    want.backward()
    assert abs(float(loss.v) - float(want.detach())) <= 1e-5 * abs(float(want.detach()))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    assert rel(z.g, zt.grad) <= 1e-4
    for k in ("conv_post.weight", "ups.0.weight_v", "resblocks.14.convs2.2.bias", "conv_pre.bias"):
        assert rel(dec.P[k].g.reshape(Pt[k].shape), Pt[k].grad) <= 1e-3, k


def test_encoder_quantizer_chain(golden_dir):
    """wav -> MelStyleEncoder / PosteriorAudioEncoder -> proj -> VQ (straight-through + commitment loss) -> nearest x2 up-sampling -> a
    random time window per clip (vq2.py:846-862): the head of SynthesizerTrn.forward on one tape, against torch.autograd through the encoder
    oracle (pinned to the REAL reference) and the quantizer formulas (core_vq.py:303-322)."""
    import numpy as np
    from oracle import encoder_oracle as EO
    from oracle import vq_mel_oracle as V
    from ttts_b200.vqvae.train_encoder import EncoderGraph
    K = TorchRefKernels()
    enc = np.load(os.path.join(golden_dir, "encoder.npz"))
    P = EO.init_params(seed=5)
    wav, lengths, eps = torch.tensor(enc["wav"]), torch.tensor(enc["lengths"]), torch.tensor(enc["eps"])
    spec = torch.tensor(V.spectrogram(enc["wav"]))
    E = torch.tensor(enc["E"])
    g0 = torch.Generator().manual_seed(21)
    R = torch.randn(3, 192, 8, generator=g0)
    starts = [3, 20, 0]
    graph = EncoderGraph(K, P)
    z, x = graph.forward(spec, wav, lengths=lengths, eps=eps)
    q, commit, codes = graph.ops.vq(x, E)
    assert np.array_equal(codes.view(3, -1).numpy(), enc["codes"][0])                     # the reference's own codes for these clips
    seg = graph.ops.slice_t(graph.ops.upsample2(q), starts, 8)
    o = graph.ops
    loss = o.add(o.scale(commit, 1.0), o.lsgan(seg, 0.0))                                 # commit + mean(seg^2) ...
    loss.g = torch.ones(1)
    seg.g = R.clone()                                                                     # ... + <seg, R>, seeded directly on the window
    graph.tape.backward()
    Pt = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    ot = EO.encode(Pt, spec, wav, lengths=lengths, eps=eps)
    xt = ot["x"]
    qt_codes = K._vq_codes(xt.detach(), E)
    qq = E[qt_codes].view(3, -1, 192).permute(0, 2, 1)
    q_st = xt + (qq - xt).detach()
    commit_t = ((qq.detach() - xt) ** 2).mean()
    up = q_st.repeat_interleave(2, dim=-1)
    seg_t = torch.stack([up[b, :, s:s + 8] for b, s in enumerate(starts)])
    (commit_t + (seg_t ** 2).mean() + (seg_t * R).sum()).backward()
    floor = 1e-6 * float(torch.sqrt(sum((v.grad ** 2).sum() for v in Pt.values() if v.grad is not None)))
    for k, v in Pt.items():
        want = v.grad if v.grad is not None else torch.zeros_like(v)
        got = graph.P[k].g.reshape(v.shape) if graph.P[k].g is not None else torch.zeros_like(v)
        assert float((got - want).norm()) <= 2e-3 * float(want.norm()) + floor, k
