"""Training-mode discriminators of the VQ-VAE-GAN step (next scope row) on the tape / op set of train_encoder.py:
`MultiPeriodDiscriminator` = DiscriminatorS (grouped Conv1d stack) + five DiscriminatorP ((5,1) Conv2d over the waveform folded to
[T/p, p]) (ttts/vqvae/vq2.py:418-551) and the adversarial losses of ttts/vqvae/losses.py:7-44 as the trainer forms them
(ttts/vqvae/train.py:372-395): discriminator step on (y, y_hat.detach()), generator step = generator_loss + feature_loss.

A (k,1) Conv2d with stride (s,1) over [B, C, R, p] is a Conv1d along R applied to each of the p columns independently, so the period
discriminators run on the 1-D convolution kernels after a [B, C, R, p] -> [B p, C, R] permutation (memory plumbing).

DRAFT, NOT YET RUN ON HARDWARE: over the torch restatement of the op contract the graph reproduces logits, losses and gradients of the REAL
reference modules (tests/test_train_disc_cpu.py vs tests/golden/disc.npz)."""
import torch
import torch.nn.functional as F

from .train_encoder import Ops, Tape, Var

PERIODS = [2, 3, 5, 7, 11]
S_CONVS = [(1, 16, 15, 1, 1, 7), (16, 64, 41, 4, 4, 20), (64, 256, 41, 4, 16, 20), (256, 1024, 41, 4, 64, 20), (1024, 1024, 41, 4, 256, 20),
           (1024, 1024, 5, 1, 1, 2)]                                     # cin, cout, kernel, stride, groups, padding   (vq2.py:498-507)
P_CONVS = [(1, 32, 3), (32, 128, 3), (128, 512, 3), (512, 1024, 3), (1024, 1024, 1)]     # cin, cout, stride; kernel 5, padding 2


class DiscriminatorGraph:
    def __init__(self, K, params, tape=None, prefix=""):
        """`tape`: share one tape between graphs to differentiate through their composition (the full step); `prefix`: the sub-module's
        prefix inside `params` (e.g. "dec."), stripped from the names the graph uses"""
        params = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
        self.K = K
        self.tape = tape if tape is not None else Tape()
        self.ops = Ops(K, self.tape)
        # (k,1) Conv2d weights [Cout, Cin, k, 1] enter as Conv1d weights [Cout, Cin, k]; weight_g [Cout,1,1,1] as [Cout,1,1]
        self.P = {k: Var(v.detach().squeeze(-1).contiguous() if v.dim() == 4 else v.detach().contiguous()) for k, v in params.items()}
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}

    def _w(self, prefix):
        return self.ops.wn(self.P[prefix + "weight_v"], self.P[prefix + "weight_g"])

    # ---- memory plumbing with its (pure data movement) backward ----
    def _fold(self, x, period):
        """[B, 1, T] -> reflect-pad to a multiple of the period -> columns as batch: [B p, 1, T/p]"""
        B, C, T = x.v.shape
        n_pad = (period - T % period) % period
        v = F.pad(x.v, (0, n_pad), "reflect") if n_pad else x.v
        R = (T + n_pad) // period
        y = Var(v.view(B, C, R, period).permute(0, 3, 1, 2).reshape(B * period, C, R).contiguous())

        def bwd():
            if y.g is None:
                return
            g = y.g.view(B, period, C, R).permute(0, 2, 3, 1).reshape(B, C, R * period)
            gx = g[..., :T].clone()
            if n_pad:                                                   # reflect: padded sample j mirrors x[T - 2 - j]
                gx[..., T - 1 - n_pad:T - 1] += g[..., T:].flip(-1)
            self.ops._acc(x, gx.contiguous())
        self.tape.record(bwd)
        return y

    def disc_s(self, x):
        o, P, pre = self.ops, self.P, "discriminators.0."
        fmap = []
        for i, (cin, cout, k, st, g, pad) in enumerate(S_CONVS):
            x = o.lrelu(o.conv(x, self._w(pre + "convs.%d." % i), P[pre + "convs.%d.bias" % i], stride=st, pad=pad, groups=g), 0.1)
            fmap.append(x)
        x = o.conv(x, self._w(pre + "conv_post."), P[pre + "conv_post.bias"], pad=1)
        fmap.append(x)
        return x, fmap

    def disc_p(self, x, d):
        o, P, pre = self.ops, self.P, "discriminators.%d." % (d + 1)
        x = self._fold(x, PERIODS[d])
        fmap = []
        for i, (cin, cout, st) in enumerate(P_CONVS):
            x = o.lrelu(o.conv(x, self._w(pre + "convs.%d." % i), P[pre + "convs.%d.bias" % i], stride=st, pad=2), 0.1)
            fmap.append(x)
        x = o.conv(x, self._w(pre + "conv_post."), P[pre + "conv_post.bias"], pad=1)
        fmap.append(x)
        return x, fmap

    def forward(self, y):
        """y [B, 1, T] (a Var when the gradient with respect to the waveform is wanted).  Returns (logits, feature maps) of the 6 discriminators."""
        y = y if isinstance(y, Var) else Var(y.contiguous())
        outs, fmaps = [], []
        for d in range(6):
            lo, fm = self.disc_s(y) if d == 0 else self.disc_p(y, d - 1)
            outs.append(lo); fmaps.append(fm)
        return outs, fmaps

    def discriminator_loss(self, real, gen):
        """losses.py:18-32"""
        o, loss = self.ops, None
        for dr, dg in zip(real, gen):
            t = o.add(o.lsgan(dr, 1.0), o.lsgan(dg, 0.0))
            loss = t if loss is None else o.add(loss, t)
        return loss

    def generator_losses(self, gen, fmap_r, fmap_g):
        """generator_loss + feature_loss (losses.py:7-15, 35-44); fmap_r enter as constants (the reference detaches them)"""
        o, lg, lf = self.ops, None, None
        for dg in gen:
            t = o.lsgan(dg, 1.0)
            lg = t if lg is None else o.add(lg, t)
        for fr, fg in zip(fmap_r, fmap_g):
            for a, b in zip(fr, fg):
                t = o.l1_mean(a.v, b)
                lf = t if lf is None else o.add(lf, t)
        return lg, o.scale(lf, 2.0)

    def backward(self, loss):
        loss.g = torch.ones_like(loss.v)
        self.tape.backward()
        return {k: (v.g.reshape(self.shapes[k]) if v.g is not None else torch.zeros(self.shapes[k], device=v.v.device)) for k, v in self.P.items()}
