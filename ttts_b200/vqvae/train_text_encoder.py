"""Training-mode prior encoder of the VQ-VAE (next scope row) on the tape / op set of train_encoder.py: `enc_p_2` =
TextEncoder(192, 192, 768, heads 2, layers 6, kernel 3) (ttts/vqvae/vq2.py:101-164) -- relative-position transformers over the up-sampled
quantized latents (3 layers) and the embedded text (6 layers), MRTE cross-attention with the style vector added (vq2.py:17-50), 3 more layers,
projection to (m_p, logs_p).  Eval-mode semantics (the reference's p = 0.1 dropouts are not drawn here yet).

DRAFT, NOT YET RUN ON HARDWARE: the graph over the torch restatement of the op contract reproduces the REAL module's outputs and all
gradients (tests/test_train_text_encoder_cpu.py vs tests/golden/text_encoder.npz); the two ops it adds -- `attn` (windowed relative-position /
cross attention) and `lnc` (channel LayerNorm) -- are csrc/text_encoder_kernels.cu, checked on the CPU emulation (tests/test_emu_text_encoder_cpu.py)."""
import torch

from .train_encoder import Ops, Tape, Var

HID, HEADS, OUT, MRTE_HEADS = 192, 2, 192, 4


class TextEncoderGraph:
    """Parameter names = the reference's `enc_p_2.*` state_dict entries without the prefix."""

    def __init__(self, K, params, tape=None, prefix=""):
        params = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
        self.K = K
        self.tape = tape if tape is not None else Tape()
        self.ops = Ops(K, self.tape)
        self.P = {k: Var(v.detach().contiguous()) for k, v in params.items()}
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}

    def encoder(self, pre, n_layers, x, mask2, lens):
        """attentions.Encoder.forward (attentions.py:66-88, g = None)"""
        o, P = self.ops, self.P
        x = o.mul_mask(x, mask2)
        for i in range(n_layers):
            a = pre + "attn_layers.%d." % i
            q = o.conv(x, P[a + "conv_q.weight"], P[a + "conv_q.bias"])
            k = o.conv(x, P[a + "conv_k.weight"], P[a + "conv_k.bias"])
            v = o.conv(x, P[a + "conv_v.weight"], P[a + "conv_v.bias"])
            y = o.attn(q, k, v, P[a + "emb_rel_k"], P[a + "emb_rel_v"], lens, lens, HEADS)
            y = o.conv(y, P[a + "conv_o.weight"], P[a + "conv_o.bias"])
            x = o.lnc(o.add(x, y), P[pre + "norm_layers_1.%d.gamma" % i], P[pre + "norm_layers_1.%d.beta" % i])
            f = pre + "ffn_layers.%d." % i
            h = o.lrelu(o.conv(o.mul_mask(x, mask2), P[f + "conv_1.weight"], P[f + "conv_1.bias"], pad=1), 0.0)          # relu
            y = o.mul_mask(o.conv(o.mul_mask(h, mask2), P[f + "conv_2.weight"], P[f + "conv_2.bias"], pad=1), mask2)
            x = o.lnc(o.add(x, y), P[pre + "norm_layers_2.%d.gamma" % i], P[pre + "norm_layers_2.%d.beta" % i])
        return o.mul_mask(x, mask2)

    def forward(self, y, y_lengths, text, text_lengths, ge):
        """y [B,192,T] Var (up-sampled quantized latents), text [B,Tt] int64, ge [B,512,1] Var -> (y, stats = [m_p | logs_p]) Vars"""
        o, P = self.ops, self.P
        dev = y.v.device
        ymask = (torch.arange(y.v.shape[2], device=dev)[None, :] < y_lengths[:, None]).float().contiguous()
        tmask = (torch.arange(text.shape[1], device=dev)[None, :] < text_lengths[:, None]).float().contiguous()
        y = self.encoder("encoder_ssl.", 3, y, ymask, y_lengths)
        t = self.encoder("encoder_text.", 6, o.embedding(P["text_embedding.weight"], text), tmask, text_lengths)
        # MRTE (vq2.py:34-50)
        ssl = o.conv(o.mul_mask(y, ymask), P["mrte.c_pre.weight"], P["mrte.c_pre.bias"])
        te = o.conv(o.mul_mask(t, tmask), P["mrte.text_pre.weight"], P["mrte.text_pre.bias"])
        c = "mrte.cross_attention."
        sm, tm = o.mul_mask(ssl, ymask), o.mul_mask(te, tmask)
        q = o.conv(sm, P[c + "conv_q.weight"], P[c + "conv_q.bias"])
        k = o.conv(tm, P[c + "conv_k.weight"], P[c + "conv_k.bias"])
        v = o.conv(tm, P[c + "conv_v.weight"], P[c + "conv_v.bias"])
        x = o.conv(o.attn(q, k, v, None, None, y_lengths, text_lengths, MRTE_HEADS), P[c + "conv_o.weight"], P[c + "conv_o.bias"])
        x = o.add_bcast(o.add(x, ssl), ge)
        y = o.conv(o.mul_mask(x, ymask), P["mrte.c_post.weight"], P["mrte.c_post.bias"])
        y = self.encoder("encoder2.", 3, y, ymask, y_lengths)
        stats = o.mul_mask(o.conv(y, P["proj.weight"], P["proj.bias"]), ymask)
        return y, stats

    def grads(self):
        return {k: (v.g.reshape(self.shapes[k]) if v.g is not None else torch.zeros(self.shapes[k], device=v.v.device)) for k, v in self.P.items()}
