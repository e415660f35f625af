"""Training-mode encode half of the VQ-VAE (next scope row, SURVEY.md 8f-1: the encoder's backward inside the VQ-VAE-GAN train step,
ttts/vqvae/train.py:330-406 -> vq2.py:843-852): MelStyleEncoder, PosteriorAudioEncoder (downsampling stack, 15 ResBlock1, anti-aliased
SnakeBeta, 16-layer WN), posterior sample, proj -- forward with the activations kept, and the gradient of every parameter.

DRAFT, NOT YET RUN ON HARDWARE.  Structure:
  * a small tape (`Tape`, `Var`, `Ops`): every op is ONE forward kernel call and records ONE closure that calls its backward kernel, so the
    backward pass is the tape in reverse -- no tracing compiler, no torch.autograd on the compute path;
  * the kernels come from a backend object `K`.  The product backend is `CudaKernels` (ctypes into libttts_b200.so: ttts_conv1d_f32,
    ttts_conv1d_bwd_input / _weight, ttts_weight_norm(_bwd), ttts_snake_aa(_bwd), ... ) and raises off-GPU: there is no CPU fallback.
    tests/ref_kernels.py holds a torch restatement of the same per-op contract; tests/test_train_encoder_cpu.py runs THIS graph over it and
    checks all 414 parameter gradients against the REAL reference's (tests/golden/encoder_grads.npz), which pins the wiring of the graph;
    the CUDA kernels are checked op by op against the same contract on the CPU emulation (tests/test_emu_*) and on the GPU.
  * slicing / concatenation / reshapes are memory plumbing and use torch tensor views.
"""
import math

import torch

HID, GIN = 192, 512
RATES, KSZ = [10, 8, 2, 2, 2], [16, 16, 8, 2, 2]


class Var:
    """a tensor on the tape and the slot its gradient accumulates into"""
    __slots__ = ("v", "g")

    def __init__(self, v):
        self.v, self.g = v, None


class Tape:
    def __init__(self):
        self.steps = []

    def record(self, fn):
        self.steps.append(fn)

    def backward(self):
        for fn in reversed(self.steps):
            fn()
        self.steps = []


class Ops:
    """The op set of the encoder: forward = one kernel call on `K`, backward = one recorded closure."""

    def __init__(self, K, tape):
        self.K, self.tape = K, tape

    def _acc(self, var, g):
        if g is None:
            return
        var.g = g if var.g is None else self.K.add(var.g, g)

    def conv(self, x, w, b=None, stride=1, dil=1, pad=0, pre_lrelu=False, need_dx=True, groups=1):
        kw = {} if groups == 1 else {"groups": groups}
        y = Var(self.K.conv_fwd(x.v, w.v, b.v if b is not None else None, stride, dil, pad, pre_lrelu, **kw))

        def bwd():
            if y.g is None:
                return
            dx, dw, db = self.K.conv_bwd(y.g, x.v, w.v, stride, dil, pad, pre_lrelu, need_dx, b is not None, **kw)
            if need_dx:
                self._acc(x, dx)
            self._acc(w, dw)
            if b is not None:
                self._acc(b, db)
        self.tape.record(bwd)
        return y

    def wn(self, v, g):
        w = Var(self.K.wn_fwd(v.v, g.v))

        def bwd():
            if w.g is None:
                return
            dv, dg = self.K.wn_bwd(w.g, v.v, g.v)
            self._acc(v, dv); self._acc(g, dg)
        self.tape.record(bwd)
        return w

    def add(self, a, b):
        c = Var(self.K.add(a.v, b.v))

        def bwd():
            self._acc(a, c.g); self._acc(b, c.g)
        self.tape.record(bwd)
        return c

    def scale(self, a, s):
        c = Var(self.K.scale(a.v, s))

        def bwd():
            if c.g is not None:
                self._acc(a, self.K.scale(c.g, s))
        self.tape.record(bwd)
        return c

    def mul_mask(self, a, mask):
        c = Var(self.K.mul_mask(a.v, mask))

        def bwd():
            if c.g is not None:
                self._acc(a, self.K.mul_mask(c.g, mask))
        self.tape.record(bwd)
        return c

    def _unary(self, name, x, *consts):
        y = Var(getattr(self.K, name + "_fwd")(x.v, *consts))

        def bwd():
            if y.g is not None:
                self._acc(x, getattr(self.K, name + "_bwd")(y.g, x.v, *consts))
        self.tape.record(bwd)
        return y

    def glu(self, raw):
        return self._unary("glu", raw)

    def lrelu(self, x, slope):
        return self._unary("lrelu", x, slope)

    def tanh(self, x):
        return self._unary("tanh", x)

    def lsgan(self, x, c):
        """mean((c - x)^2) as a [1] tensor (losses.py:18-44)"""
        y = Var(self.K.lsgan_fwd(x.v, c))

        def bwd():
            if y.g is not None:
                self._acc(x, self.K.lsgan_bwd(y.g, x.v, c))
        self.tape.record(bwd)
        return y

    def attn(self, q, k, v, emb_k, emb_v, q_len, k_len, heads):
        """multi-head attention, channel-major, with an optional relative-position window (emb_k / emb_v Vars or None)"""
        ek, ev = (emb_k.v, emb_v.v) if emb_k is not None else (None, None)
        o = Var(self.K.attn_fwd(q.v, k.v, v.v, ek, ev, q_len, k_len, heads))

        def bwd():
            if o.g is None:
                return
            dq, dk, dv, dek, dev = self.K.attn_bwd(o.g, q.v, k.v, v.v, ek, ev, q_len, k_len, heads)
            self._acc(q, dq); self._acc(k, dk); self._acc(v, dv)
            if emb_k is not None:
                self._acc(emb_k, dek); self._acc(emb_v, dev)
        self.tape.record(bwd)
        return o

    def lnc(self, x, gamma, beta):
        """LayerNorm over the channel axis of [B, C, T]"""
        y = Var(self.K.lnc_fwd(x.v, gamma.v, beta.v))

        def bwd():
            if y.g is None:
                return
            dx, dg, db = self.K.lnc_bwd(y.g, x.v, gamma.v, beta.v)
            self._acc(x, dx); self._acc(gamma, dg); self._acc(beta, db)
        self.tape.record(bwd)
        return y

    def embedding(self, table, idx):
        """table [V, C] Var, idx [B, T] int64 -> [B, C, T]; memory gather, the backward scatter-adds rows"""
        y = Var(table.v[idx].transpose(1, 2).contiguous())

        def bwd():
            if y.g is not None:
                g = torch.zeros_like(table.v)
                g.index_add_(0, idx.reshape(-1), y.g.transpose(1, 2).reshape(-1, table.v.shape[1]))
                self._acc(table, g)
        self.tape.record(bwd)
        return y

    def vq(self, x, embed):
        """EuclideanCodebook lookup in training mode (core_vq.py:174-182, 303-322): returns (straight-through quantized Var, commit loss Var
        [1], codes).  The codebook itself has no gradient (EMA-updated buffers); its update is the caller's (ttts_vq_ema_update)."""
        qv, cv, codes = self.K.vq_fwd(x.v, embed)
        q, c = Var(qv), Var(cv)

        def bwd():
            if q.g is None and c.g is None:
                return
            self._acc(x, self.K.vq_bwd(q.g, c.g, x.v, embed, codes))
        self.tape.record(bwd)
        return q, c, codes

    def upsample2(self, x):
        """F.interpolate(x, size = 2 T, mode="nearest") (vq2.py:853-855): memory plumbing; the backward sums the pairs"""
        y = Var(x.v.repeat_interleave(2, dim=-1).contiguous())

        def bwd():
            if y.g is not None:
                B, C, T2 = y.g.shape
                self._acc(x, self.K.add(y.g[..., 0::2].contiguous(), y.g[..., 1::2].contiguous()))
        self.tape.record(bwd)
        return y

    def slice_t(self, x, starts, size):
        """commons.slice_segments (vq2.py:860-862): per-sequence windows [start, start + size) along time; backward scatters"""
        B = x.v.shape[0]
        y = Var(torch.stack([x.v[b, :, int(starts[b]):int(starts[b]) + size] for b in range(B)]).contiguous())

        def bwd():
            if y.g is not None:
                g = torch.zeros_like(x.v)
                for b in range(B):
                    g[b, :, int(starts[b]):int(starts[b]) + size] = y.g[b]
                self._acc(x, g)
        self.tape.record(bwd)
        return y

    def logmel(self, wav):
        """mel_spectrogram_torch of the v2 front end (data_utils.py:106-156) of a waveform Var [B, L]: the mel-reconstruction loss
        differentiates through it (train.py:357-366,389)"""
        y = Var(self.K.logmel_fwd(wav.v))

        def bwd():
            if y.g is not None:
                self._acc(wav, self.K.logmel_bwd(y.g, wav.v))
        self.tape.record(bwd)
        return y

    def kl(self, z_p, logs_q, m_p, logs_p, mask):
        """kl_loss of the trainer (losses.py:47-61) as a [1] tensor; all four inputs are Vars, mask [B, T]"""
        y = Var(self.K.kl_fwd(z_p.v, logs_q.v, m_p.v, logs_p.v, mask))

        def bwd():
            if y.g is None:
                return
            gs = self.K.kl_bwd(y.g, z_p.v, logs_q.v, m_p.v, logs_p.v, mask)
            for var, g in zip((z_p, logs_q, m_p, logs_p), gs):
                self._acc(var, g)
        self.tape.record(bwd)
        return y

    def l1_mean(self, a_const, b):
        """mean(|a - b|) with a detached (losses.py:7-15)"""
        y = Var(self.K.l1_fwd(a_const, b.v))

        def bwd():
            if y.g is not None:
                self._acc(b, self.K.l1_bwd(y.g, a_const, b.v))
        self.tape.record(bwd)
        return y

    def add_bcast(self, x, c):
        """x [B,C,T] + c [B,C,1]"""
        y = Var(self.K.add_bcast_fwd(x.v, c.v))

        def bwd():
            if y.g is not None:
                self._acc(x, y.g); self._acc(c, self.K.add_bcast_bwd(y.g))
        self.tape.record(bwd)
        return y

    def convT(self, x, w, b, stride, pad):
        """ConvTranspose1d with weight [Cin, Cout, K] (Generator.ups, vq2.py:369-378)"""
        y = Var(self.K.convT_fwd(x.v, w.v, b.v if b is not None else None, stride, pad))

        def bwd():
            if y.g is None:
                return
            dx, dw, db = self.K.convT_bwd(y.g, x.v, w.v, stride, pad, b is not None)
            self._acc(x, dx); self._acc(w, dw)
            if b is not None:
                self._acc(b, db)
        self.tape.record(bwd)
        return y

    def mish(self, x):
        return self._unary("mish", x)

    def gate(self, raw, cond):
        y = Var(self.K.gate_fwd(raw.v, cond.v if cond is not None else None))

        def bwd():
            if y.g is None:
                return
            draw, dcond = self.K.gate_bwd(y.g, raw.v, cond.v if cond is not None else None)
            self._acc(raw, draw)
            if cond is not None:
                self._acc(cond, dcond)
        self.tape.record(bwd)
        return y

    def snake(self, x, log_alpha, log_beta, filt):
        y = Var(self.K.snake_fwd(x.v, log_alpha.v, log_beta.v, filt))

        def bwd():
            if y.g is None:
                return
            dx, dla, dlb = self.K.snake_bwd(y.g, x.v, log_alpha.v, log_beta.v, filt)
            self._acc(x, dx); self._acc(log_alpha, dla); self._acc(log_beta, dlb)
        self.tape.record(bwd)
        return y

    def mha(self, q, k, v, lens, heads, temperature):
        o = Var(self.K.mha_fwd(q.v, k.v, v.v, lens, heads, temperature))

        def bwd():
            if o.g is None:
                return
            dq, dk, dv = self.K.mha_bwd(o.g, q.v, k.v, v.v, lens, heads, temperature)
            self._acc(q, dq); self._acc(k, dk); self._acc(v, dv)
        self.tape.record(bwd)
        return o

    def masked_mean(self, x, lens):
        T = x.v.shape[-1]
        y = Var(self.K.masked_mean_fwd(x.v, lens))                  # [B, C]

        def bwd():
            if y.g is not None:
                self._acc(x, self.K.masked_mean_bwd(y.g, lens, T))
        self.tape.record(bwd)
        return y

    def posterior(self, stats, eps, mask):
        z = Var(self.K.posterior_fwd(stats.v, eps, mask))

        def bwd():
            if z.g is not None:
                self._acc(stats, self.K.posterior_bwd(z.g, stats.v, eps, mask))
        self.tape.record(bwd)
        return z

    # ---- memory plumbing (views / copies, no arithmetic) ----
    def slice_c(self, x, c0, c1):
        y = Var(x.v[:, c0:c1].contiguous())

        def bwd():
            if y.g is None:
                return
            g = torch.zeros_like(x.v)
            g[:, c0:c1] = y.g
            self._acc(x, g)
        self.tape.record(bwd)
        return y

    def cat_c(self, a, b):
        ca = a.v.shape[1]
        y = Var(torch.cat([a.v, b.v], dim=1))

        def bwd():
            if y.g is not None:
                self._acc(a, y.g[:, :ca].contiguous()); self._acc(b, y.g[:, ca:].contiguous())
        self.tape.record(bwd)
        return y

    def flip_c(self, x):
        y = Var(torch.flip(x.v, [1]).contiguous())

        def bwd():
            if y.g is not None:
                self._acc(x, torch.flip(y.g, [1]).contiguous())
        self.tape.record(bwd)
        return y

    def reshape(self, x, shape):
        y = Var(x.v.reshape(shape))

        def bwd():
            if y.g is not None:
                self._acc(x, y.g.reshape(x.v.shape))
        self.tape.record(bwd)
        return y


def kaiser_sinc_filter12(device):
    """alias_free_torch/filter.py: kaiser_sinc_filter1d(cutoff 0.25, half_width 0.3, kernel 12) -- a constant buffer of the reference"""
    cutoff, half_width, ks = 0.25, 0.3, 12
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7)
    window = torch.kaiser_window(ks, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (f / f.sum()).to(device=device, dtype=torch.float32).contiguous()


def wn_stack(o, P, pre, x, mask2, g, hid, n_layers, kernel_size=5):
    """WN (modules.py:136-221; dilation_rate 1): gated convolutions conditioned on g, residual and skip paths.  Used by the posterior
    encoders (16 layers) and by the flow's coupling layers (4 layers)."""
    wn = lambda prefix: o.wn(P[prefix + "weight_v"], P[prefix + "weight_g"])
    out = None
    gc = o.reshape(o.conv(g, wn(pre + "cond_layer."), P[pre + "cond_layer.bias"]), (g.v.shape[0], 2 * hid * n_layers))
    for i in range(n_layers):
        raw = o.conv(x, wn(pre + "in_layers.%d." % i), P[pre + "in_layers.%d.bias" % i], pad=(kernel_size - 1) // 2)
        acts = o.gate(raw, o.slice_c(gc, i * 2 * hid, (i + 1) * 2 * hid))
        rs = o.conv(acts, wn(pre + "res_skip_layers.%d." % i), P[pre + "res_skip_layers.%d.bias" % i])
        if i < n_layers - 1:
            x = o.mul_mask(o.add(x, o.slice_c(rs, 0, hid)), mask2)
            skip = o.slice_c(rs, hid, 2 * hid)
        else:
            skip = rs
        out = skip if out is None else o.add(out, skip)
    return o.mul_mask(out, mask2)


class EncoderGraph:
    """The training graph over the reference's parameter names (the state_dict of ref_enc.*, enc_p.*, proj.*)."""

    def __init__(self, K, params, tape=None, prefix=""):
        """`tape`: share one tape between graphs to differentiate through their composition (the full step); `prefix`: the sub-module's
        prefix inside `params` (e.g. "dec."), stripped from the names the graph uses"""
        params = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
        self.K = K
        self.tape = tape if tape is not None else Tape()
        self.ops = Ops(K, self.tape)
        # leaves; 2-D Linear weights enter as [out, in, 1] convolution weights
        self.P = {k: Var(v.detach().unsqueeze(-1).contiguous() if v.dim() == 2 else v.detach().contiguous()) for k, v in params.items()}
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}

    def _wn(self, prefix, old=True):
        if old:
            return self.ops.wn(self.P[prefix + "weight_v"], self.P[prefix + "weight_g"])
        return self.ops.wn(self.P[prefix + "parametrizations.weight.original1"], self.P[prefix + "parametrizations.weight.original0"])

    def mel_style_encoder(self, x, mask2, lens):
        """modules.py:686-764.  x [B,1025,T] already masked (a constant: no gradient flows to the spectrogram)."""
        o, P, pre = self.ops, self.P, "ref_enc."
        h = o.mish(o.conv(x, P[pre + "spectral.0.fc.weight"], P[pre + "spectral.0.fc.bias"], need_dx=False))
        h = o.mish(o.conv(h, P[pre + "spectral.3.fc.weight"], P[pre + "spectral.3.fc.bias"]))
        for i in range(2):
            c = o.conv(h, P[pre + "temporal.%d.conv1.conv.weight" % i], P[pre + "temporal.%d.conv1.conv.bias" % i], pad=2)
            h = o.add(h, o.glu(c))
        h = o.mul_mask(h, mask2)
        q = o.conv(h, P[pre + "slf_attn.w_qs.weight"], P[pre + "slf_attn.w_qs.bias"])
        k = o.conv(h, P[pre + "slf_attn.w_ks.weight"], P[pre + "slf_attn.w_ks.bias"])
        v = o.conv(h, P[pre + "slf_attn.w_vs.weight"], P[pre + "slf_attn.w_vs.bias"])
        att = o.mha(q, k, v, lens, 2, math.sqrt(128.0))
        h = o.add(o.conv(att, P[pre + "slf_attn.fc.weight"], P[pre + "slf_attn.fc.bias"]), h)
        h = o.conv(h, P[pre + "fc.fc.weight"], P[pre + "fc.fc.bias"])
        return o.reshape(o.masked_mean(h, lens), (h.v.shape[0], GIN, 1))

    def resblock1(self, prefix, x, k):
        o, P = self.ops, self.P
        for t, d in enumerate((1, 3, 5)):
            xt = o.conv(x, self._wn(prefix + "convs1.%d." % t, old=False), P[prefix + "convs1.%d.bias" % t], dil=d, pad=(k * d - d) // 2, pre_lrelu=True)
            xt = o.conv(xt, self._wn(prefix + "convs2.%d." % t, old=False), P[prefix + "convs2.%d.bias" % t], pad=(k - 1) // 2, pre_lrelu=True)
            x = o.add(xt, x)
        return x

    def wn_stack(self, x, mask2, g, pre="enc_p."):
        return wn_stack(self.ops, self.P, pre + "enc.", x, mask2, g, HID, 16)

    def posterior_audio_encoder(self, spec, wav, mask2, g, eps, pre="enc_p."):
        """PosteriorAudioEncoder (vq2.py:667-745); `pre` = "enc_p." or "enc_q." (the same class, vq2.py:814-828)"""
        o, P = self.ops, self.P
        a = o.conv(wav, P[pre + "down_pre.weight"], P[pre + "down_pre.bias"], pad=3, need_dx=False)
        for i in range(5):
            a = o.conv(a, self._wn(pre + "downs.%d." % i), P[pre + "downs.%d.bias" % i], stride=RATES[i], pad=(KSZ[i] - 1) // 2)
            xs = None
            for j, k in enumerate((3, 7, 11)):
                r = self.resblock1(pre + "resblocks.%d." % (i * 3 + j), a, k)
                xs = r if xs is None else o.add(xs, r)
            a = o.scale(xs, 1.0 / 3.0)
        a = o.snake(a, P[pre + "activation_post.act.alpha"], P[pre + "activation_post.act.beta"], kaiser_sinc_filter12(wav.v.device))
        a = o.mul_mask(o.conv(a, P[pre + "conv_post.weight"], P[pre + "conv_post.bias"], pad=3), mask2)
        x = o.mul_mask(o.conv(spec, P[pre + "pre.weight"], P[pre + "pre.bias"], need_dx=False), mask2)
        x = self.wn_stack(x, mask2, g, pre)
        stats = o.mul_mask(o.conv(o.cat_c(x, a), P[pre + "proj.weight"], P[pre + "proj.bias"]), mask2)
        return o.posterior(stats, eps, mask2), stats

    def forward(self, spec, wav, lengths=None, eps=None):
        """spec [B,1025,T] (constant), wav [B,L].  Returns Vars z [B,192,T] and x [B,192,T/2] (the input of the quantizer)."""
        B, _, T = spec.shape
        dev = spec.device
        if lengths is None:
            lengths = torch.full((B,), T, dtype=torch.int64, device=dev)
        mask2 = (torch.arange(T, device=dev)[None, :] < lengths[:, None]).float().contiguous()
        specm = Var(self.K.mul_mask(spec.contiguous(), mask2))
        ge = self.mel_style_encoder(specm, mask2, lengths)
        z, stats = self.posterior_audio_encoder(Var(spec.contiguous()), Var(wav.unsqueeze(1).contiguous()), mask2, ge, eps)
        x = self.ops.conv(z, self.P["proj.weight"], self.P["proj.bias"], stride=2)
        self.out = dict(ge=ge, z=z, stats=stats, x=x)
        return z, x

    def backward(self, dz=None, dx=None):
        """Seed the gradients of z and / or x and run the tape.  Returns name -> gradient with the parameter's own shape."""
        if dz is not None:
            self.out["z"].g = dz.contiguous()
        if dx is not None:
            self.out["x"].g = dx.contiguous()
        self.tape.backward()
        return {k: (v.g.reshape(self.shapes[k]) if v.g is not None else torch.zeros(self.shapes[k], device=v.v.device)) for k, v in self.P.items()}


def dgrad_by_phase(conv1d_fn, dy, w, Tin, stride, pad):
    """Input gradient of a STRIDED convolution (dilation 1) as `stride` independent stride-1 convolutions, one per phase of the input
    position: ti + pad = stride q + phi only meets the taps k = phi + stride j, so
        dx[ci, stride q + phi - pad] = sum_co sum_j w[co, ci, phi + stride j] dy[co, q - j]
    -- a "full" correlation of dy with the phase's sub-kernel (K / stride taps).  conv1d_dgrad_kernel walks ALL K taps for every position and
    multiplies zeros for (stride - 1) / stride of them (the down-sampling stack has stride 10 / kernel 16: 90 % waste); the phases together
    do exactly the useful FLOPs, on the pipelined forward kernels.  ConvTranspose1d's forward is the same operation (Generator.ups).
    conv1d_fn(x, w, pad) = stride-1 cross-correlation; dy [B,Cout,Tout], w [Cout,Cin,K] -> dx [B,Cin,Tin]."""
    B, Cout, Tout = dy.shape
    Cin, K = w.shape[1], w.shape[2]
    dx = torch.zeros(B, Cin, Tin, dtype=dy.dtype, device=dy.device)
    for phi in range(min(stride, K)):
        J = (K - phi + stride - 1) // stride                        # taps phi, phi + stride, ... < K
        wt = w[:, :, phi::stride].flip(2).transpose(0, 1).contiguous()          # [Cin, Cout, J], tap order reversed
        out = conv1d_fn(dy, wt, J - 1)                                # [B, Cin, Tout + J - 1]: out[q] = sum_j w_phi[j] dy[q - j]
        q_min = max(0, -((phi - pad) // stride))                     # first q with stride q + phi - pad >= 0
        t0 = stride * q_min + phi - pad
        if t0 >= Tin:
            continue
        n = min((Tin - 1 - t0) // stride + 1, out.shape[2] - q_min)
        if n > 0:
            dx[:, :, t0:t0 + stride * (n - 1) + 1:stride] = out[:, :, q_min:q_min + n]
    return dx


class CudaKernels:
    """The product backend: every method is one call (conv_bwd: two) into libttts_b200.so on the current stream.  Device tensors only."""

    def __init__(self, lib=None):
        """`lib`: None = libttts_b200.so (the product).  tests/emu_kernels.py passes the host builds of the same sources instead, to run THIS
        class's argument marshalling on the CPU emulation."""
        from .. import _lib as L
        from . import encoder as E
        self.L, self.E = L, E
        if lib is None:
            lib = L.lib()
            E._protos(lib)
        self.lib = lib
        self._train_protos(lib)
        import os
        self.dgrad_as_forward = os.environ.get("TTTS_DGRAD_FWD", "1") != "0"
        # wide convolutions (both channel counts multiples of 64 and >= 128, >= 4096 output positions) as split-bf16 GEMMs on the tcgen05 GEMM
        self.conv_gemm = os.environ.get("TTTS_TRAIN_GEMM", "1") != "0" and os.environ.get("TTTS_DIFF_TC", "1") != "0"
        # dil = 1 layers: all taps in one GEMM reduction over an overlapped view of the activation rows (TTTS_GEMM_TAPCAT=0: one GEMM pair per tap)
        self.tap_concat = os.environ.get("TTTS_GEMM_TAPCAT", "1") != "0"
        self.wgrad_concat = self.tap_concat and os.environ.get("TTTS_GEMM_WCAT", "1") != "0"        # the same for the weight gradient
        # 32- / 64-channel layers too (TTTS_GEMM_NARROW=0: fp32 kernels).  r2ae, B = 64: the fp32 families lose 83 ms, the GEMM family gains
        # 44 ms and the layout conversions around such thin GEMMs ate the rest (549 vs 541 ms); with the conversions on 64 x 64 tiles / packed
        # stores and the 16-byte-load bias gradient (r2ai) the route wins: 501 vs 521 ms per step
        self.gemm_narrow = os.environ.get("TTTS_GEMM_NARROW", "1") != "0"
        # dilated stride-1 layers as ordinary convolutions over de-interleaved sub-clips (TTTS_GEMM_DEINT=0: one GEMM pair per tap / fp32 kernels)
        self.deinterleave = os.environ.get("TTTS_GEMM_DEINT", "1") != "0"
        self.fused_wprep = os.environ.get("TTTS_GEMM_WPREP", "1") != "0"          # one kernel for the routed layers' weight operands (0: torch ops)

    @staticmethod
    def _train_protos(lib):
        import ctypes
        if not getattr(lib, "_train_protos", False):
            vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
            lib.ttts_conv1d_bwd_input.argtypes = [vp, vp, vp, vp] + [i32] * 10 + [vp]
            lib.ttts_conv1d_bwd_weight.argtypes = [vp, vp, vp, vp] + [i32] * 9 + [vp]
            lib.ttts_stft_mel_bwd.argtypes = [vp, i32, i32, i32, i32, i32, vp, f32, i32, vp, vp, vp, f32, vp, i32, vp, vp]
            lib.ttts_ew_add.argtypes = [vp, vp, vp, i64, vp]
            lib.ttts_ew_scale.argtypes = [vp, f32, vp, i64, vp]
            lib.ttts_ew_mul_mask.argtypes = [vp, vp, vp, i32, i32, i32, vp]
            lib.ttts_glu.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
            lib.ttts_mish.argtypes = [vp, vp, vp, i64, i32, vp]
            lib.ttts_wn_gate.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
            lib.ttts_weight_norm_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
            lib.ttts_lrelu.argtypes = [vp, vp, vp, i64, f32, i32, vp]
            lib.ttts_tanh.argtypes = [vp, vp, vp, i64, i32, vp]
            lib.ttts_add_bcast.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
            lib.ttts_sum_t.argtypes = [vp, vp, i32, i32, vp]
            lib.ttts_bias_grad.argtypes = [vp, vp, i32, i32, i32, vp]
            lib.ttts_lsgan_loss.argtypes = [vp, f32, i64, vp, vp, vp]
            lib.ttts_lsgan_loss_bwd.argtypes = [vp, f32, vp, i64, vp, vp]
            lib.ttts_l1_mean.argtypes = [vp, vp, i64, vp, vp, vp]
            lib.ttts_l1_mean_bwd.argtypes = [vp, vp, vp, i64, vp, vp]
            lib.ttts_gconv1d.argtypes = [vp] * 4 + [i32] * 8 + [vp]
            lib.ttts_gconv1d_bwd.argtypes = [vp] * 5 + [i32] * 8 + [vp]
            lib.ttts_attn_small.argtypes = [vp] * 8 + [i32] * 6 + [vp]
            lib.ttts_attn_small_bwd.argtypes = [vp] * 13 + [i32] * 6 + [vp]
            lib.ttts_layernorm_c.argtypes = [vp] * 5 + [i32] * 3 + [vp]
            lib.ttts_layernorm_c_bwd.argtypes = [vp] * 8 + [i32] * 3 + [vp]
            lib.ttts_kl_loss.argtypes = [vp] * 5 + [i32, i32, i32, vp, vp, vp]
            lib.ttts_kl_loss_bwd.argtypes = [vp] * 6 + [i32, i32, i32, vp, vp, vp, vp, vp]
            lib.ttts_snake_aa_bwd.argtypes = [vp] * 8 + [i32, i32, i32, vp]
            lib.ttts_mha_small_bwd.argtypes = [vp] * 8 + [i32, i32, i32, i32, f32, vp]
            lib.ttts_masked_mean_bwd.argtypes = [vp, vp, vp, i32, i32, i32, vp]
            lib.ttts_posterior_sample_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
            try:
                lib.ttts_cl_split.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
                lib.ttts_cl_unpack.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]
                lib.ttts_conv_w_concat.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
            except AttributeError:
                pass                                                  # an emulation set without csrc/diffusion_kernels.cu
            lib._train_protos = True

    def _st(self):
        return self.L.stream_ptr().value

    def _chk(self, rc, what):
        self.L.check(rc, what)

    @staticmethod
    def _p(t):
        return t.data_ptr() if t is not None else None

    def _device_check(self, ts):
        self.L.require_cuda(*ts)

    def _req(self, *ts):
        self._device_check([t for t in ts if t is not None])
        for t in ts:
            assert t is None or (t.is_contiguous() and t.dtype in (torch.float32, torch.int64)), "contiguous fp32 / int64 tensors only"


    # ---------------------------------------------------------------- wide convolutions on the tcgen05 GEMM ----------------------------------------------------------------
    # x = hi + lo, w = hi + lo (bf16 each), x w ~ hi hi + lo hi + hi lo with fp32 accumulation in TMEM: fp32-grade results (2e-5 relative against an
    # fp64 convolution, tests/test_gpu_diffusion.py) at tensor-core speed.  ttts_cl_split writes the activation as position-major [hi | lo] rows,
    # sample t of clip b at row b Tp + pad + t of a zero buffer (Tp = stride * ceil((T + 2 pad) / stride)), so that output row m = b Tp / stride + to
    # reads tap k at row stride m + k dil: an A operand with ROW STRIDE `stride`, no im2col, no tap reads a neighbouring clip.  Then
    #   forward : D[m, co]  = sum_k [xh | xl](s m + k dil) . [wh_k | wh_k]^T  +  xh(s m + k dil) . wl_k^T                 (2 launches per tap)
    #   dgrad   : Dx[s m + k dil, ci] += [dyh | dyl](m) . [wh_k ; wh_k]  +  dyh(m) . wl_k          (fp32 red.add into a strided view of a zero buffer)
    #   wgrad   : dW_k      = dyh^T . [xh | xl](s . + k dil)  (two column blocks, summed)  +  dyl^T . xh(s . + k dil)        (split-K, fp32 red.add)
    # r2o / r2p: the diffusion step 353 -> 196 ms; the period discriminators' 512 -> 1024 (stride 3) and 1024 -> 1024 layers are ~60 % of the
    # VQ-VAE-GAN step's FLOPs.
    GEMM_MIN_POSITIONS = 4096

    def _gemm_ok(self, x, w, stride, dil, pad, groups):
        if not self.conv_gemm or not x.is_cuda or groups != 1 or not hasattr(self.lib, "ttts_cl_split"):
            return False
        return self._gemm_shape_ok(x, w, stride, dil, pad)

    def _gemm_shape_ok(self, x, w, stride, dil, pad):
        """the shape rule of the route (device-independent: tests/test_gemm_conv_route_cpu.py drives whole training graphs through it on CPU)"""
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        Tout = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
        if not (K <= 16 and stride <= 8 and Tout >= 1 and B * Tout >= self.GEMM_MIN_POSITIONS):
            return False
        if Cin % 64 == 0 and Cout % 64 == 0 and min(Cin, Cout) >= 128:
            return True
        # narrow layers (the Generator's 64- and 32-channel ResBlock convolutions on 5 120 / 10 240 positions per clip): only in the
        # tap-concatenated form, where every product reduces over K 2 Cin >= 64 and the accumulator is swept twice instead of 2 K times
        # (per tap these layers were bound by that traffic and lost to the fp32 kernels)
        return (self.gemm_narrow and self.tap_concat and self.wgrad_concat and (dil == 1 or self._deinterleaved(T, K, stride, dil, pad)) and K > 1
                and Cin % 32 == 0 and Cout % 32 == 0 and pad <= dil * (K - 1) and (stride == 1 or Cout % 64 == 0))

    def release_buffers(self):
        """drop the transient buffers of the GEMM route (position-major activations, accumulators, weight operands: one set per distinct
        layer shape, a few GB at B = 64; they are kept between steps so that a step allocates nothing)"""
        self.__dict__.pop("_gemm_pool", None)

    def _buf(self, key, shape, dtype, dev, zero=False):
        """transient buffers, one per key: every use is ordered on the current stream (never freed: see release_buffers)"""
        pool = self.__dict__.setdefault("_gemm_pool", {})
        key = (key, tuple(shape), dtype, dev)
        if key not in pool:
            pool[key] = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
        return pool[key]

    def _cl_split(self, tag, x, rows_per_clip=None, row_off=1, rows=None, lrelu=False, dil=1):
        """x [B,C,T] -> [rows, 2C] bf16 rows [hi | lo]; the rows the kernel never writes are the zero padding (default geometry: one zero row
        before every clip and one after the last).  dil > 1: de-interleaved, B dil sub-clips of T / dil positions (include/ttts_b200.h)"""
        B, C, T = x.shape
        rows_per_clip = T + 1 if rows_per_clip is None else rows_per_clip
        rows = 2 + B * rows_per_clip if rows is None else rows
        buf = self._buf((tag, B, C, T, rows_per_clip, row_off, dil), (rows, 2 * C), torch.bfloat16, x.device, zero=True)
        self._chk(self.lib.ttts_cl_split(self._p(x), self._p(buf), B, C, T, rows_per_clip, row_off, int(bool(lrelu)), dil, self._st()), "ttts_cl_split")
        return buf

    def _cl_unpack(self, D, B, C, T, rows_per_clip=None, row_off=0, lrelu_x=None, dil=1):
        """D (position-major fp32 rows) -> [B,C,T]; lrelu_x: multiply by leaky_relu'(lrelu_x) on the way (the input gradient of a layer that
        applies the activation to its input)"""
        y = torch.empty(B, C, T, dtype=torch.float32, device=D.device)
        rows_per_clip = T + 1 if rows_per_clip is None else rows_per_clip
        self._chk(self.lib.ttts_cl_unpack(self._p(D), self._p(y), B, C, T, D.stride(0), rows_per_clip, row_off, self._p(lrelu_x), dil, self._st()), "ttts_cl_unpack")
        return y

    @staticmethod
    def _split_weights(w):
        wk = w.permute(2, 0, 1).contiguous()                          # [K, Cout, Cin]
        wh = wk.bfloat16()
        wl = (wk - wh.float()).bfloat16()
        return wh, wl

    def _concat_weights(self, w, flip_t=False):
        """w [Cout, Cin, K] -> the two B operands of the tap-concatenated GEMM pair, [R, K 2 Cr] bf16 each: taps side by side, per tap
        [wh | wh] and [wl | 0].  flip_t: the input-gradient form (rows = input channels, taps reversed).  One kernel (ttts_conv_w_concat);
        the buffers are per shape and reused in stream order"""
        Cout, Cin, K = w.shape
        if self.fused_wprep and hasattr(self.lib, "ttts_conv_w_concat"):
            R, Cr = (Cin, Cout) if flip_t else (Cout, Cin)
            W1 = self._buf(("W1", flip_t), (R, K * 2 * Cr), torch.bfloat16, w.device)
            W2 = self._buf(("W2", flip_t), (R, K * 2 * Cr), torch.bfloat16, w.device)
            self._chk(self.lib.ttts_conv_w_concat(self._p(w), self._p(W1), self._p(W2), Cout, Cin, K, int(flip_t), self._st()), "ttts_conv_w_concat")
            return W1, W2
        if flip_t:
            w = w.flip(2).transpose(0, 1).contiguous()
            Cout, Cin, K = w.shape
        wk = w.permute(0, 2, 1)                                       # [Cout, K, Cin] (a view)
        wh = wk.bfloat16()                                            # contiguous [Cout, K, Cin]
        wl = (wk - wh).bfloat16()                                     # bf16 promotes to fp32 in the subtraction
        Z = self._buf("wzero", (Cout, K, Cin), torch.bfloat16, w.device, zero=True)
        return torch.cat([wh, wh], dim=2).view(Cout, K * 2 * Cin), torch.cat([wl, Z], dim=2).view(Cout, K * 2 * Cin)

    @staticmethod
    def _gemm_geometry(B, T, K, stride, dil, pad):
        Tout = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
        Tout_p = (T + 2 * pad + stride - 1) // stride              # output rows per clip (>= Tout; the extra ones are never read back)
        Tp = stride * Tout_p                                          # input rows per clip (>= T + 2 pad)
        return Tout, Tout_p, Tp, B * Tout_p, B * Tp + dil * (K - 1) + 1

    def _gemm_conv_fwd(self, x, w, b, stride, dil, pad, pre_lrelu, lrelu_x=None, split_tag="x", w_flip_t=False):
        """lrelu_x: see _cl_unpack (used when this is the input gradient of another layer); w_flip_t: w is the weight of the layer whose INPUT
        gradient this is ([x channels out, x channels in, K] read transposed with the taps reversed); returns (y, X) so that a caller can
        reuse the split buffer -- plain callers take [0]"""
        self._req(x, w, b)
        B, Cin, T = x.shape
        K = w.shape[2]
        d = dil if self._deinterleaved(T, K, stride, dil, pad) else 1
        Tout, Tout_p, Tp, M, rows = self._gemm_geometry(B * d, (T + d - 1) // d, K, stride, dil // d, pad // d)
        X = self._cl_split(split_tag, x, Tp, pad // d, rows, pre_lrelu, d)
        return self._gemm_conv_core(X, B, Cin, T, w, b, stride, dil, pad, lrelu_x, d, w_flip_t), X

    def _deinterleaved(self, T, K, stride, dil, pad):
        """a dilated stride-1 layer whose padding is a multiple of the dilation runs as an ordinary convolution over B dil sub-clips (one per
        residue class of the time index, ceil(T / dil) positions each -- the shorter ones end in zero rows; the layout kernels write / read
        that order), i.e. in the tap-concatenated form"""
        return dil > 1 and K > 1 and stride == 1 and pad % dil == 0 and self.tap_concat and self.deinterleave

    def _gemm_conv_core(self, X, B, Cin, T, w, b, stride, dil, pad, lrelu_x=None, d=1, w_flip_t=False):
        """X = the split activation; d > 1: de-interleaved (then B, T, dil, pad are still the layer's own; the GEMMs see B d clips of T / d)"""
        L = self.L
        K = w.shape[2]
        Cout = w.shape[1] if w_flip_t else w.shape[0]
        Tout_full = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
        B_out, T_out = B, Tout_full
        B, T, dil, pad = B * d, (T + d - 1) // d, dil // d, pad // d
        Tout, Tout_p, Tp, M, rows = self._gemm_geometry(B, T, K, stride, dil, pad)
        assert Tout == (T_out + d - 1) // d
        D = self._buf("D", (M, Cout), torch.float32, X.device)
        bias = b.clone() if (b is not None and b.data_ptr() % 16) else b
        if dil == 1 and K > 1 and self.tap_concat:
            # all taps in ONE reduction: row m of the A operand is the K consecutive buffer rows stride m .. stride m + K - 1 seen as one
            # K 2C-wide row -- an OVERLAPPED view (row pitch stride 2C < row length K 2C; the TMA tensor map takes the pitch as given), so
            # still no im2col.  D is written once and updated once instead of 2 K read-modify-write sweeps (r2aa launch list: the per-tap
            # GEMMs of a 128-channel K = 11 layer were bound by that fp32 accumulator traffic, not by the tensor cores).  The hi . lo
            # product reads the same operand against [wl_k | 0] (the zero half costs MMA time, which is not what bounds these layers).
            A = X.as_strided((M, K * 2 * Cin), (stride * 2 * Cin, 1))
            W1, W2 = self._concat_weights(w, w_flip_t)
            L.gemm(A, W1, D, epi=L.EPI_F32, bias=bias)
            L.gemm(A, W2, D, epi=L.EPI_F32_ADD)
            return self._cl_unpack(D, B_out, Cout, T_out, Tout_p, 0, lrelu_x, d)
        wh, wl = self._split_weights(w.flip(2).transpose(0, 1).contiguous() if w_flip_t else w)
        first = True
        for k in range(K):
            A = X[k * dil:k * dil + stride * (M - 1) + 1:stride]
            L.gemm(A, torch.cat([wh[k], wh[k]], dim=1), D, epi=L.EPI_F32 if first else L.EPI_F32_ADD, bias=bias if first else None)
            L.gemm(A[:, :Cin], wl[k], D, epi=L.EPI_F32_ADD)
            first = False
        return self._cl_unpack(D, B_out, Cout, T_out, Tout_p, 0, lrelu_x, d)

    def _gemm_conv_bwd(self, dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db):
        L = self.L
        dy = dy.contiguous()
        self._req(dy, x, w)
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        Tout_full = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
        # a de-interleaved dilated layer: from here on B d sub-clips of T / d positions, dilation 1 (the layout kernels are told d)
        d = dil if self._deinterleaved(T, K, stride, dil, pad) else 1
        Be, Te, dile, pade = B * d, (T + d - 1) // d, dil // d, pad // d
        Tout, Tout_p, Tp, M, rows = self._gemm_geometry(Be, Te, K, stride, dile, pade)
        dx, DY = None, None
        if need_dx and stride == 1 and (dil == 1 or d > 1) and K > 1 and pad <= dil * (K - 1) and self.tap_concat:
            # input gradient of a stride-1 convolution = the forward route on dy with the taps flipped and the channel roles swapped
            # (padding dil (K - 1) - pad): one reduction over all taps instead of 2 K red.add sweeps over a zeroed buffer; the derivative of
            # a leaky ReLU on the layer's input is applied by the conversion back to [B, C, T]
            dx, DYp = self._gemm_conv_fwd(dy, w, None, 1, dil, dil * (K - 1) - pad, False, lrelu_x=x if pre_lrelu else None, split_tag="dyp",
                                          w_flip_t=True)
            assert dx.shape == x.shape
            if 2 * pade == K - 1:
                # a "same" convolution: the padded rows of dy just written have the clip pitch the weight gradient needs (T + K - 1), shifted
                # by the padding -- one split of dy serves both gradients
                DY = DYp[K - 1 - pade:K - 1 - pade + M]
            need_dx = False
        if DY is None:
            DY = self._cl_split("dy", dy, Tout_p, 0, M, False, d)
        dil, pad = dile, pade                                         # the GEMMs below see the effective layer
        if need_dx:
            wh, wl = self._split_weights(w)
            Dx = self._buf("Dx", (rows, Cin), torch.float32, x.device)
            Dx.zero_()
            for k in range(K):
                O = Dx[k * dil:k * dil + stride * (M - 1) + 1:stride]
                L.gemm(DY, torch.cat([wh[k], wh[k]], dim=0), O, b_mn=True, epi=L.EPI_F32_ADD)
                L.gemm(DY[:, :Cout], wl[k], O, b_mn=True, epi=L.EPI_F32_ADD)
            dx = self._cl_unpack(Dx, B, Cin, T, Tp, pad, x if pre_lrelu else None, d)
        X = self._cl_split("x", x, Tp, pad, rows, pre_lrelu, d)
        if dil == 1 and K > 1 and self.wgrad_concat:
            # all taps as column blocks of ONE product: the overlapped view again, now as the MN-major B operand [positions, K 2 Cin]; dy is read
            # once per product instead of once per tap.  Both halves of dy meet the whole view, so the lo . lo term comes along for free.
            Xov = X.as_strided((M, K * 2 * Cin), (stride * 2 * Cin, 1))
            acc = torch.zeros(Cout, K * 2 * Cin, dtype=torch.float32, device=x.device)
            L.gemm(DY[:, :Cout], Xov, acc, a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=16)
            L.gemm(DY[:, Cout:], Xov, acc, a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=16)
            acc = acc.view(Cout, K, 2 * Cin)
            dw = (acc[:, :, :Cin] + acc[:, :, Cin:]).permute(0, 2, 1).contiguous()
        else:
            acc = torch.zeros(K, Cout, 2 * Cin, dtype=torch.float32, device=x.device)
            for k in range(K):
                Xk = X[k * dil:k * dil + stride * (M - 1) + 1:stride]
                L.gemm(DY[:, :Cout], Xk, acc[k], a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=16)
                L.gemm(DY[:, Cout:], Xk[:, :Cin], acc[k][:, :Cin], a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=16)
            dw = (acc[:, :, :Cin] + acc[:, :, Cin:]).permute(1, 2, 0).contiguous()
        db = None
        if need_db:
            db = torch.zeros(Cout, dtype=torch.float32, device=x.device)
            self._chk(self.lib.ttts_bias_grad(self._p(dy), self._p(db), B, Cout, Tout_full, self._st()), "ttts_bias_grad")
        return dx, dw, db

    # ---- convolution / weight norm ----
    def conv_fwd(self, x, w, b, stride, dil, pad, pre_lrelu, groups=1):
        if groups > 1:                                                 # DiscriminatorS (vq2.py:498-507): csrc/conv1d_grouped.cu
            assert dil == 1 and not pre_lrelu
            self._req(x, w, b)
            B, Cin, Tin = x.shape
            Cout, _, K = w.shape
            y = torch.empty(B, Cout, (Tin + 2 * pad - K) // stride + 1, dtype=torch.float32, device=x.device)
            self._chk(self.lib.ttts_gconv1d(self._p(x), self._p(w), self._p(b), self._p(y), B, Cin, Tin, Cout, K, stride, pad, groups, self._st()), "ttts_gconv1d")
            return y
        if self._gemm_ok(x, w, stride, dil, pad, groups):
            return self._gemm_conv_fwd(x, w, b, stride, dil, pad, pre_lrelu)[0]
        return self.E.conv1d(x, w, b, stride=stride, dil=dil, pad=pad, pre_lrelu=pre_lrelu)

    def conv_bwd(self, dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db, groups=1):
        if groups > 1:
            assert dil == 1 and not pre_lrelu
            self._req(dy, x, w)
            B, Cin, Tin = x.shape
            Cout, _, K = w.shape
            dx = torch.empty_like(x) if need_dx else None
            dw = torch.zeros_like(w)
            self._chk(self.lib.ttts_gconv1d_bwd(self._p(dy), self._p(x), self._p(w), self._p(dx), self._p(dw), B, Cin, Tin, Cout, K, stride, pad, groups,
                                                self._st()), "ttts_gconv1d_bwd")
            db = None
            if need_db:
                db = torch.zeros(Cout, dtype=torch.float32, device=x.device)
                self._chk(self.lib.ttts_bias_grad(self._p(dy), self._p(db), B, Cout, dy.shape[-1], self._st()), "ttts_bias_grad")
            return dx, dw, db
        if self._gemm_ok(x, w, stride, dil, pad, groups):
            return self._gemm_conv_bwd(dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db)
        self._req(dy, x, w)
        B, Cin, Tin = x.shape
        Cout, _, K = w.shape
        p, lib, st = self._p, self.lib, self._st()
        dx = None
        pad_t = dil * (K - 1) - pad
        if need_dx and stride == 1 and pad_t >= 0 and self.dgrad_as_forward and hasattr(self.E, "conv1d") and dy.is_cuda:
            # the input gradient of a stride-1 convolution IS a convolution: dx = conv(dy, w^T flipped in k, padding dil (K - 1) - pad).  The
            # forward implicit-GEMM kernels (cp.async pipeline, 64-channel tiles) run it ~2x faster than conv1d_dgrad_kernel (r2m launch list);
            # exact fp32 either way, the summation order differs.
            wt = w.flip(2).transpose(0, 1).contiguous()
            dx = self.E.conv1d(dy, wt, None, stride=1, dil=dil, pad=pad_t, tc=False)
            assert dx.shape == x.shape
            if pre_lrelu:
                self._chk(lib.ttts_lrelu(p(x), p(dx), p(dx), x.numel(), 0.1, 1, st), "ttts_lrelu (dgrad)")
        elif need_dx and stride > 1 and dil == 1 and self.dgrad_as_forward and hasattr(self.E, "conv1d") and dy.is_cuda:
            dx = dgrad_by_phase(self._phase_conv, dy, w, Tin, stride, pad)
            if pre_lrelu:
                self._chk(lib.ttts_lrelu(p(x), p(dx), p(dx), x.numel(), 0.1, 1, st), "ttts_lrelu (dgrad)")
        elif need_dx:
            dx = torch.empty_like(x)
            self._chk(lib.ttts_conv1d_bwd_input(p(dy), p(w), p(x), p(dx), B, Cin, Tin, Cout, K, stride, dil, pad, int(pre_lrelu), 0, st), "ttts_conv1d_bwd_input")
        dw = torch.zeros_like(w)
        db = torch.zeros(Cout, dtype=torch.float32, device=x.device) if need_db else None
        self._chk(lib.ttts_conv1d_bwd_weight(p(dy), p(x), p(dw), p(db), B, Cin, Tin, Cout, K, stride, dil, pad, int(pre_lrelu), st), "ttts_conv1d_bwd_weight")
        return dx, dw, db

    def _phase_conv(self, a, wt, pd):
        """stride-1 cross-correlation of one phase (dgrad_by_phase): the GEMM route when the phase's sub-kernel qualifies, else the fp32 kernels"""
        if self._gemm_ok(a, wt, 1, 1, pd, 1):
            return self._gemm_conv_fwd(a, wt, None, 1, 1, pd, False)[0]
        return self.E.conv1d(a, wt, None, stride=1, dil=1, pad=pd, tc=False)

    def convT_fwd(self, x, w, b, stride, pad):
        """conv_transpose1d = the input gradient of a convolution whose weight is w read as [Cout' = Cin, Cin' = Cout, K]"""
        self._req(x, w, b)
        B, Cin, T = x.shape
        _, Cout, K = w.shape
        Tout = (T - 1) * stride - 2 * pad + K
        p, lib, st = self._p, self.lib, self._st()
        if self.dgrad_as_forward and hasattr(self.E, "conv1d") and x.is_cuda:
            y = dgrad_by_phase(self._phase_conv, x, w, Tout, stride, pad)
        else:
            y = torch.empty(B, Cout, Tout, dtype=torch.float32, device=x.device)
            self._chk(lib.ttts_conv1d_bwd_input(p(x), p(w), None, p(y), B, Cout, Tout, Cin, K, stride, 1, pad, 0, 0, st), "ttts_conv1d_bwd_input (convT forward)")
        if b is not None:
            self._chk(lib.ttts_add_bcast(p(y), p(b), p(y), B, Cout, Tout, 0, st), "ttts_add_bcast (bias)")
        return y

    def convT_bwd(self, dy, x, w, stride, pad, need_db):
        self._req(dy, x, w)
        B, Cin, T = x.shape
        _, Cout, K = w.shape
        Tout = dy.shape[-1]
        p, lib, st = self._p, self.lib, self._st()
        dx = self.conv_fwd(dy, w, None, stride, 1, pad, False)                              # [B, Cin, T]
        assert dx.shape == x.shape
        if self._gemm_ok(dy, w, stride, 1, pad, 1):
            # the weight gradient of the convolution dy -> x whose input gradient this layer's forward is: roles of x and dy swapped
            _, dw, _ = self._gemm_conv_bwd(x, dy, w, stride, 1, pad, False, False, False)
        else:
            dw = torch.zeros_like(w)
            self._chk(lib.ttts_conv1d_bwd_weight(p(x), p(dy), p(dw), None, B, Cout, Tout, Cin, K, stride, 1, pad, 0, st), "ttts_conv1d_bwd_weight (convT)")
        db = None
        if need_db:
            db = torch.zeros(Cout, dtype=torch.float32, device=x.device)
            self._chk(lib.ttts_bias_grad(p(dy), p(db), B, Cout, Tout, st), "ttts_bias_grad")
        return dx, dw, db

    def _scratch(self, dev):
        s = getattr(self, "_red", None)
        if s is None or s.device != dev:
            s = self._red = torch.empty(256, dtype=torch.float32, device=dev)
        return s

    def lsgan_fwd(self, x, c):
        self._req(x)
        o = torch.empty(1, dtype=torch.float32, device=x.device)
        self._chk(self.lib.ttts_lsgan_loss(self._p(x), float(c), x.numel(), self._p(self._scratch(x.device)), self._p(o), self._st()), "ttts_lsgan_loss")
        return o

    def lsgan_bwd(self, dL, x, c):
        self._req(dL, x)
        d = torch.empty_like(x)
        self._chk(self.lib.ttts_lsgan_loss_bwd(self._p(x), float(c), self._p(dL), x.numel(), self._p(d), self._st()), "ttts_lsgan_loss_bwd")
        return d

    def attn_fwd(self, q, k, v, emb_k, emb_v, q_len, k_len, heads):
        self._req(q, k, v, emb_k, emb_v, q_len, k_len)
        B, C, Tq = q.shape
        win = (emb_k.shape[1] - 1) // 2 if emb_k is not None else 0
        o = torch.empty_like(q)
        self._chk(self.lib.ttts_attn_small(self._p(q), self._p(k), self._p(v), self._p(emb_k), self._p(emb_v), self._p(q_len), self._p(k_len), self._p(o),
                                           B, C, Tq, k.shape[2], heads, win, self._st()), "ttts_attn_small")
        return o

    def attn_bwd(self, do, q, k, v, emb_k, emb_v, q_len, k_len, heads):
        do = do.contiguous()
        self._req(do, q, k, v, emb_k, emb_v, q_len, k_len)
        B, C, Tq = q.shape
        win = (emb_k.shape[1] - 1) // 2 if emb_k is not None else 0
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dek = torch.zeros_like(emb_k) if emb_k is not None else None
        dev = torch.zeros_like(emb_v) if emb_v is not None else None
        self._chk(self.lib.ttts_attn_small_bwd(self._p(do), self._p(q), self._p(k), self._p(v), self._p(emb_k), self._p(emb_v), self._p(q_len), self._p(k_len),
                                               self._p(dq), self._p(dk), self._p(dv), self._p(dek), self._p(dev), B, C, Tq, k.shape[2], heads, win,
                                               self._st()), "ttts_attn_small_bwd")
        return dq, dk, dv, dek, dev

    def lnc_fwd(self, x, gamma, beta):
        self._req(x, gamma, beta)
        B, C, T = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(B * T * 2, dtype=torch.float32, device=x.device)
        self._chk(self.lib.ttts_layernorm_c(self._p(x), self._p(gamma), self._p(beta), self._p(y), self._p(stats), B, C, T, self._st()), "ttts_layernorm_c")
        return y

    def lnc_bwd(self, dy, x, gamma, beta):
        dy = dy.contiguous()
        self._req(dy, x, gamma, beta)
        B, C, T = x.shape
        # the statistics are recomputed (2 small launches) rather than kept per call site
        stats = torch.empty(B * T * 2, dtype=torch.float32, device=x.device)
        tmp = torch.empty_like(x)
        self._chk(self.lib.ttts_layernorm_c(self._p(x), self._p(gamma), self._p(beta), self._p(tmp), self._p(stats), B, C, T, self._st()), "ttts_layernorm_c")
        dx, dg, db = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(beta)
        scratch = torch.empty(B * T * 2, dtype=torch.float32, device=x.device)
        self._chk(self.lib.ttts_layernorm_c_bwd(self._p(dy), self._p(x), self._p(stats), self._p(gamma), self._p(dx), self._p(dg), self._p(db),
                                                self._p(scratch), B, C, T, self._st()), "ttts_layernorm_c_bwd")
        return dx, dg, db

    def vq_fwd(self, x, embed):
        """`embed`: the codebook tensor [K, D], or the reference-shaped EuclideanCodebook module of ttts_b200.vqvae.quantize -- then the step is
        the module's training forward (core_vq.py:205-230): lookup -> expire_codes_ -> EMA update of cluster_size / embed_avg / embed"""
        from . import quantize as Q
        cb = None
        if not torch.is_tensor(embed):
            cb, embed = embed, embed.embed
        self._req(x, embed)
        hist = embed_sum = None
        if cb is not None:
            B, D, Nn = x.shape
            # first training batch only: k-means initialisation (core_vq.py:209; kmeans_init=True is the quantizer's default, so a fresh
            # module has an all-zero codebook until this runs -- without it every vector would map to code 0 and the EMA would collapse)
            cb.init_embed_(x.detach().permute(0, 2, 1).reshape(B * Nn, D))
            embed = cb.embed
            hist = torch.zeros(cb.codebook_size, dtype=torch.float32, device=x.device)
            embed_sum = torch.zeros_like(cb.embed)
            embed = embed.clone()                                     # the backward gathers from the PRE-update codebook
        codes, q, commit = Q.vq_lookup(x, embed, True, True, straight_through=True, want_commit=True, hist=hist, embed_sum=embed_sum)
        if cb is not None:
            B, D, Nn = x.shape
            cb.expire_codes_(x.detach().permute(0, 2, 1).reshape(B * Nn, D))
            cb._ema_update(hist, embed_sum)
            self._vq_embed = embed
        return q, commit.reshape(1), codes

    def vq_bwd(self, dq, dcommit, x, embed, codes):
        from . import quantize as Q
        if not torch.is_tensor(embed):
            embed = self._vq_embed
        lib = self.lib
        Q._protos(lib)
        B, D, Nn = x.shape
        dx = torch.empty_like(x)
        dq = dq.contiguous() if dq is not None else None
        dc = dcommit.contiguous() if dcommit is not None else None
        self._chk(lib.ttts_vq_backward(self._p(x), B, D, Nn, 1, self._p(embed), self._p(codes), self._p(dq), self._p(dc), self._p(dx), self._st()),
                  "ttts_vq_backward")
        return dx

    def logmel_fwd(self, wav):
        from . import mel as M
        self._req(wav)
        return M.mel_spectrogram_torch(wav, 2048, 128, 32000, 640, 2048, 0, None)

    def logmel_bwd(self, dmel, wav):
        from . import mel as M
        dmel = dmel.contiguous()
        self._req(dmel, wav)
        B, Lw = wav.shape
        window, _ = M._stft_consts(2048, 2048, wav.device)
        lo, off, w = M._mel_consts("slaney", 32000, 2048, 128, 0, None, wav.device)
        dwav = torch.zeros_like(wav)
        self._chk(self.lib.ttts_stft_mel_bwd(self._p(wav), B, Lw, 2048, 640, 704, self._p(window), 1e-6, 128, lo.data_ptr(), off.data_ptr(), w.data_ptr(),
                                             1e-5, self._p(dmel), dmel.shape[-1], self._p(dwav), self._st()), "ttts_stft_mel_bwd")
        return dwav

    def kl_fwd(self, z_p, logs_q, m_p, logs_p, mask):
        self._req(z_p, logs_q, m_p, logs_p, mask)
        B, C, T = z_p.shape
        dev = z_p.device
        if getattr(self, "_red2", None) is None or self._red2.device != dev:
            self._red2 = torch.empty(512, dtype=torch.float32, device=dev)
        self._kl_out = torch.empty(2, dtype=torch.float32, device=dev)
        self._chk(self.lib.ttts_kl_loss(self._p(z_p), self._p(logs_q), self._p(m_p), self._p(logs_p), self._p(mask), B, C, T, self._p(self._red2),
                                        self._p(self._kl_out), self._st()), "ttts_kl_loss")
        return self._kl_out[:1]

    def kl_bwd(self, dL, z_p, logs_q, m_p, logs_p, mask):
        self._req(z_p, logs_q, m_p, logs_p, mask)
        B, C, T = z_p.shape
        gs = [torch.empty_like(z_p) for _ in range(4)]
        dL = dL.contiguous()
        self._chk(self.lib.ttts_kl_loss_bwd(self._p(z_p), self._p(m_p), self._p(logs_p), self._p(mask), self._p(dL), self._p(self._kl_out), B, C, T,
                                            self._p(gs[0]), self._p(gs[1]), self._p(gs[2]), self._p(gs[3]), self._st()), "ttts_kl_loss_bwd")
        return gs

    def l1_fwd(self, a, b):
        self._req(a, b)
        o = torch.empty(1, dtype=torch.float32, device=b.device)
        self._chk(self.lib.ttts_l1_mean(self._p(a), self._p(b), b.numel(), self._p(self._scratch(b.device)), self._p(o), self._st()), "ttts_l1_mean")
        return o

    def l1_bwd(self, dL, a, b):
        self._req(dL, a, b)
        d = torch.empty_like(b)
        self._chk(self.lib.ttts_l1_mean_bwd(self._p(a), self._p(b), self._p(dL), b.numel(), self._p(d), self._st()), "ttts_l1_mean_bwd")
        return d

    def lrelu_fwd(self, x, slope):
        self._req(x)
        y = torch.empty_like(x)
        self._chk(self.lib.ttts_lrelu(self._p(x), None, self._p(y), x.numel(), float(slope), 0, self._st()), "ttts_lrelu")
        return y

    def lrelu_bwd(self, dy, x, slope):
        self._req(dy, x)
        d = torch.empty_like(x)
        self._chk(self.lib.ttts_lrelu(self._p(x), self._p(dy), self._p(d), x.numel(), float(slope), 1, self._st()), "ttts_lrelu (backward)")
        return d

    def tanh_fwd(self, x):
        self._req(x)
        y = torch.empty_like(x)
        self._chk(self.lib.ttts_tanh(self._p(x), None, self._p(y), x.numel(), 0, self._st()), "ttts_tanh")
        return y

    def tanh_bwd(self, dy, x):
        self._req(dy, x)
        d = torch.empty_like(x)
        self._chk(self.lib.ttts_tanh(self._p(x), self._p(dy), self._p(d), x.numel(), 1, self._st()), "ttts_tanh (backward)")
        return d

    def add_bcast_fwd(self, x, c):
        c = c.contiguous()
        self._req(x, c)
        B, C, T = x.shape
        y = torch.empty_like(x)
        self._chk(self.lib.ttts_add_bcast(self._p(x), self._p(c), self._p(y), B, C, T, 1, self._st()), "ttts_add_bcast")
        return y

    def add_bcast_bwd(self, dy):
        self._req(dy)
        B, C, T = dy.shape
        o = torch.empty(B, C, 1, dtype=torch.float32, device=dy.device)
        self._chk(self.lib.ttts_sum_t(self._p(dy), self._p(o), B * C, T, self._st()), "ttts_sum_t")
        return o

    def wn_fwd(self, v, g):
        self._req(v, g)
        w = torch.empty_like(v)
        self._chk(self.lib.ttts_weight_norm(self._p(v), self._p(g), self._p(w), v.shape[0], v[0].numel(), self._st()), "ttts_weight_norm")
        return w

    def wn_bwd(self, dw, v, g):
        self._req(dw, v, g)
        dv, dg = torch.empty_like(v), torch.empty_like(g)
        self._chk(self.lib.ttts_weight_norm_bwd(self._p(dw), self._p(v), self._p(g), self._p(dv), self._p(dg), v.shape[0], v[0].numel(), self._st()),
                  "ttts_weight_norm_bwd")
        return dv, dg

    # ---- element-wise ----
    def add(self, a, b):
        self._req(a, b)
        o = torch.empty_like(a)
        self._chk(self.lib.ttts_ew_add(self._p(a), self._p(b), self._p(o), a.numel(), self._st()), "ttts_ew_add")
        return o

    def scale(self, a, s):
        self._req(a)
        o = torch.empty_like(a)
        self._chk(self.lib.ttts_ew_scale(self._p(a), float(s), self._p(o), a.numel(), self._st()), "ttts_ew_scale")
        return o

    def mul_mask(self, a, mask):
        self._req(a, mask)
        B, C, T = a.shape
        o = torch.empty_like(a)
        self._chk(self.lib.ttts_ew_mul_mask(self._p(a), self._p(mask), self._p(o), B, C, T, self._st()), "ttts_ew_mul_mask")
        return o

    def glu_fwd(self, raw):
        self._req(raw)
        B, C2, T = raw.shape
        y = torch.empty(B, C2 // 2, T, dtype=torch.float32, device=raw.device)
        self._chk(self.lib.ttts_glu(self._p(raw), None, self._p(y), B, C2 // 2, T, 0, self._st()), "ttts_glu")
        return y

    def glu_bwd(self, dy, raw):
        self._req(dy, raw)
        B, C2, T = raw.shape
        d = torch.empty_like(raw)
        self._chk(self.lib.ttts_glu(self._p(raw), self._p(dy), self._p(d), B, C2 // 2, T, 1, self._st()), "ttts_glu (backward)")
        return d

    def mish_fwd(self, x):
        self._req(x)
        y = torch.empty_like(x)
        self._chk(self.lib.ttts_mish(self._p(x), None, self._p(y), x.numel(), 0, self._st()), "ttts_mish")
        return y

    def mish_bwd(self, dy, x):
        self._req(dy, x)
        d = torch.empty_like(x)
        self._chk(self.lib.ttts_mish(self._p(x), self._p(dy), self._p(d), x.numel(), 1, self._st()), "ttts_mish (backward)")
        return d

    def gate_fwd(self, raw, cond):
        self._req(raw, cond)
        B, H2, T = raw.shape
        y = torch.empty(B, H2 // 2, T, dtype=torch.float32, device=raw.device)
        self._chk(self.lib.ttts_wn_gate(self._p(raw), self._p(cond), None, self._p(y), None, B, H2 // 2, T, 0, self._st()), "ttts_wn_gate")
        return y

    def gate_bwd(self, dy, raw, cond):
        self._req(dy, raw, cond)
        B, H2, T = raw.shape
        d = torch.empty_like(raw)
        dc = torch.empty_like(cond) if cond is not None else None
        self._chk(self.lib.ttts_wn_gate(self._p(raw), self._p(cond), self._p(dy), self._p(d), self._p(dc), B, H2 // 2, T, 1, self._st()), "ttts_wn_gate (backward)")
        return d, dc

    # ---- SnakeBeta, attention, mean, posterior ----
    def snake_fwd(self, x, la, lb, filt):
        self._req(x, la, lb, filt)
        B, C, T = x.shape
        y = torch.empty_like(x)
        self._chk(self.lib.ttts_snake_aa(self._p(x), self._p(la), self._p(lb), self._p(filt), self._p(y), B, C, T, self._st()), "ttts_snake_aa")
        return y

    def snake_bwd(self, dy, x, la, lb, filt):
        self._req(dy, x, la, lb, filt)
        B, C, T = x.shape
        dx, dla, dlb = torch.empty_like(x), torch.zeros_like(la), torch.zeros_like(lb)
        self._chk(self.lib.ttts_snake_aa_bwd(self._p(dy), self._p(x), self._p(la), self._p(lb), self._p(filt), self._p(dx), self._p(dla), self._p(dlb),
                                             B, C, T, self._st()), "ttts_snake_aa_bwd")
        return dx, dla, dlb

    def mha_fwd(self, q, k, v, lens, heads, temperature):
        self._req(q, k, v, lens)
        B, C, T = q.shape
        o = torch.empty_like(q)
        self._chk(self.lib.ttts_mha_small(self._p(q), self._p(k), self._p(v), self._p(lens), self._p(o), B, C, T, heads, float(temperature), self._st()),
                  "ttts_mha_small")
        return o

    def mha_bwd(self, do, q, k, v, lens, heads, temperature):
        self._req(do, q, k, v, lens)
        B, C, T = q.shape
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        self._chk(self.lib.ttts_mha_small_bwd(self._p(do), self._p(q), self._p(k), self._p(v), self._p(lens), self._p(dq), self._p(dk), self._p(dv),
                                              B, C, T, heads, float(temperature), self._st()), "ttts_mha_small_bwd")
        return dq, dk, dv

    def masked_mean_fwd(self, x, lens):
        self._req(x, lens)
        B, C, T = x.shape
        y = torch.empty(B, C, dtype=torch.float32, device=x.device)
        self._chk(self.lib.ttts_masked_mean(self._p(x), self._p(lens), self._p(y), B, C, T, self._st()), "ttts_masked_mean")
        return y

    def masked_mean_bwd(self, dy, lens, T):
        dy = dy.contiguous()
        self._req(dy, lens)
        B, C = dy.shape
        dx = torch.empty(B, C, T, dtype=torch.float32, device=dy.device)
        self._chk(self.lib.ttts_masked_mean_bwd(self._p(dy), self._p(lens), self._p(dx), B, C, T, self._st()), "ttts_masked_mean_bwd")
        return dx

    def posterior_fwd(self, stats, eps, mask):
        self._req(stats, eps, mask)
        B, C2, T = stats.shape
        z = torch.empty(B, C2 // 2, T, dtype=torch.float32, device=stats.device)
        self._chk(self.lib.ttts_posterior_sample(self._p(stats), self._p(eps), self._p(mask), self._p(z), B, C2 // 2, T, self._st()), "ttts_posterior_sample")
        return z

    def posterior_bwd(self, dz, stats, eps, mask):
        self._req(dz, stats, eps, mask)
        B, C2, T = stats.shape
        d = torch.empty_like(stats)
        self._chk(self.lib.ttts_posterior_sample_bwd(self._p(dz), self._p(stats), self._p(eps), self._p(mask), self._p(d), B, C2 // 2, T, self._st()),
                  "ttts_posterior_sample_bwd")
        return d
