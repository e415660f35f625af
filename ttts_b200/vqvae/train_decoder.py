"""Training-mode decoder of the VQ-VAE (next scope row): the HiFi-GAN-style `Generator` of SynthesizerTrn
(ttts/vqvae/vq2.py:341-416, built at :798-807) on the tape / op set of train_encoder.py -- conv_pre + cond(g), five
[leaky_relu(0.1) -> weight-normed ConvTranspose1d -> mean of three ResBlock1], leaky_relu(0.01), conv_post (no bias), tanh.
ConvTranspose1d needs no kernel of its own: its forward IS the convolution's input-gradient kernel and its two gradients are the
convolution's forward and weight-gradient kernels with the roles of input and output swapped (include/ttts_b200.h: ttts_bias_grad).

DRAFT, NOT YET RUN ON HARDWARE: over the torch restatement of the op contract (tests/ref_kernels.py) the graph reproduces the waveform and
all parameter gradients of the REAL reference Generator (tests/test_train_decoder_cpu.py vs tests/golden/decoder.npz)."""
import torch

from .train_encoder import Ops, Tape, Var

RATES, KSZ, RES_K = [10, 8, 2, 2, 2], [16, 16, 8, 2, 2], (3, 7, 11)


class DecoderGraph:
    """Parameter names = the reference's `dec.*` state_dict entries without the prefix."""

    def __init__(self, K, params, tape=None, prefix=""):
        """`tape`: share one tape between graphs to differentiate through their composition (the full step); `prefix`: the sub-module's
        prefix inside `params` (e.g. "dec."), stripped from the names the graph uses"""
        params = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
        self.K = K
        self.tape = tape if tape is not None else Tape()
        self.ops = Ops(K, self.tape)
        self.P = {k: Var(v.detach().contiguous()) for k, v in params.items()}
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}

    def resblock1(self, prefix, x, k):
        o, P = self.ops, self.P
        for t, d in enumerate((1, 3, 5)):
            w1 = o.wn(P[prefix + "convs1.%d.parametrizations.weight.original1" % t], P[prefix + "convs1.%d.parametrizations.weight.original0" % t])
            w2 = o.wn(P[prefix + "convs2.%d.parametrizations.weight.original1" % t], P[prefix + "convs2.%d.parametrizations.weight.original0" % t])
            xt = o.conv(x, w1, P[prefix + "convs1.%d.bias" % t], dil=d, pad=(k * d - d) // 2, pre_lrelu=True)
            xt = o.conv(xt, w2, P[prefix + "convs2.%d.bias" % t], pad=(k - 1) // 2, pre_lrelu=True)
            x = o.add(xt, x)
        return x

    def forward(self, z, g=None):
        """z [B, 192, T] latent (a Var so that its gradient is available: it feeds the encoder / flow side), g [B, 512, 1] Var or None."""
        o, P = self.ops, self.P
        self.z = z if isinstance(z, Var) else Var(z.contiguous())
        x = o.conv(self.z, P["conv_pre.weight"], P["conv_pre.bias"], pad=3)
        if g is not None:
            self.g = g if isinstance(g, Var) else Var(g.contiguous())
            x = o.add_bcast(x, o.conv(self.g, P["cond.weight"], P["cond.bias"]))
        for i, (u, k) in enumerate(zip(RATES, KSZ)):
            x = o.lrelu(x, 0.1)
            w = o.wn(P["ups.%d.weight_v" % i], P["ups.%d.weight_g" % i])
            x = o.convT(x, w, P["ups.%d.bias" % i], u, (k - u) // 2)
            xs = None
            for j, rk in enumerate(RES_K):
                r = self.resblock1("resblocks.%d." % (i * 3 + j), x, rk)
                xs = r if xs is None else o.add(xs, r)
            x = o.scale(xs, 1.0 / len(RES_K))
        x = o.lrelu(x, 0.01)
        x = o.conv(x, P["conv_post.weight"], None, pad=3)
        self.y = o.tanh(x)
        return self.y

    def backward(self, dy):
        self.y.g = dy.contiguous()
        self.tape.backward()
        return {k: (v.g.reshape(self.shapes[k]) if v.g is not None else torch.zeros(self.shapes[k], device=v.v.device)) for k, v in self.P.items()}
