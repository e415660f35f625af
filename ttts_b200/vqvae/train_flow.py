"""Training-mode flow of the VQ-VAE (next scope row) on the tape / op set of train_encoder.py:
`ResidualCouplingBlock(192, 192, 5, 1, 4, gin_channels=512)` in the forward direction (ttts/vqvae/vq2.py:209-246, used at :858) -- four
mean-only coupling layers (modules.py:405-459: x1 <- post(WN(pre(x0) mask, g)) mask + x1 mask) each followed by a channel flip -- and the
KL term it feeds (losses.py:47-61).  DRAFT, NOT YET RUN ON HARDWARE: over the torch restatement of the op contract it reproduces z_p, the
loss and all gradients of the REAL reference module (tests/test_train_flow_cpu.py vs tests/golden/flow.npz)."""
import torch

from .train_encoder import Ops, Tape, Var, wn_stack

CH, HID, NL, NF = 192, 192, 4, 4


class FlowGraph:
    """Parameter names = the reference's `flow.*` state_dict entries without the prefix."""

    def __init__(self, K, params, tape=None, prefix=""):
        params = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
        self.K = K
        self.tape = tape if tape is not None else Tape()
        self.ops = Ops(K, self.tape)
        self.P = {k: Var(v.detach().contiguous()) for k, v in params.items()}
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}

    def forward(self, z, mask2, g):
        """z [B,192,T] Var, mask2 [B,T], g [B,512,1] Var -> z_p"""
        o, P = self.ops, self.P
        x = z
        for f in range(NF):
            p = "flows.%d." % (2 * f)
            x0, x1 = o.slice_c(x, 0, CH // 2), o.slice_c(x, CH // 2, CH)
            h = o.mul_mask(o.conv(x0, P[p + "pre.weight"], P[p + "pre.bias"]), mask2)
            h = wn_stack(o, P, p + "enc.", h, mask2, g, HID, NL)
            m = o.mul_mask(o.conv(h, P[p + "post.weight"], P[p + "post.bias"]), mask2)
            x = o.flip_c(o.cat_c(x0, o.add(m, o.mul_mask(x1, mask2))))
        return x

    def backward(self, loss):
        loss.g = torch.ones_like(loss.v)
        self.tape.backward()
        return {k: (v.g.reshape(self.shapes[k]) if v.g is not None else torch.zeros(self.shapes[k], device=v.v.device)) for k, v in self.P.items()}
