"""TEST INFRASTRUCTURE: the per-op contract of the encoder's training kernels (ttts_b200/vqvae/train_encoder.py: `K`) restated in plain torch on
CPU.  Forward functions follow oracle/encoder_oracle.py (pinned to the REAL reference); every backward is the local torch.autograd of that
forward, so this file is the specification the CUDA kernels are checked against op by op (CPU emulation and GPU), and running the training
graph over it pins the graph's wiring against the reference's gradients (tests/test_train_encoder_cpu.py).  Never imported by the product."""
import torch
import torch.nn.functional as F


def _vjp(fn, inputs, dy):
    xs = [t.detach().clone().requires_grad_(True) for t in inputs]
    y = fn(*xs)
    return torch.autograd.grad(y, xs, dy, allow_unused=True)


class TorchRefKernels:
    # ---- convolution: y = conv1d(leaky_relu(x, 0.1) if pre_lrelu else x, w, b)            ttts_conv1d_f32 (post = 0) / ttts_conv1d_bwd_* ----
    @staticmethod
    def _conv(x, w, b, stride, dil, pad, pre_lrelu, groups=1):
        return F.conv1d(F.leaky_relu(x, 0.1) if pre_lrelu else x, w, b, stride=stride, dilation=dil, padding=pad, groups=groups)

    def conv_fwd(self, x, w, b, stride, dil, pad, pre_lrelu, groups=1):
        return self._conv(x, w, b, stride, dil, pad, pre_lrelu, groups)

    def conv_bwd(self, dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db, groups=1):
        dx, dw = _vjp(lambda a, c: self._conv(a, c, None, stride, dil, pad, pre_lrelu, groups), [x, w], dy)
        return (dx if need_dx else None), dw, (dy.sum(dim=(0, 2)) if need_db else None)

    # ---- log-mel spectrogram of the v2 front end: mel_spectrogram_torch(y, 2048, 128, 32000, 640, 2048, 0, None)   (data_utils.py:106-156) ----
    @staticmethod
    def _logmel(wav):
        from oracle import vq_mel_oracle as V
        basis = torch.tensor(V.mel_basis_slaney(32000, 2048, 128, 0.0, None).astype("float32"))
        xp = F.pad(wav.unsqueeze(1), (704, 704), mode="reflect").squeeze(1)
        spec = torch.stft(xp, 2048, hop_length=640, win_length=2048, window=torch.hann_window(2048), center=False, onesided=True, return_complex=True)
        mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-6)
        return torch.log(torch.clamp(basis @ mag, min=1e-5))

    def logmel_fwd(self, wav):                                       # wav [B, L] -> [B, 128, L / 640]
        return self._logmel(wav)

    def logmel_bwd(self, dmel, wav):
        return _vjp(self._logmel, [wav], dmel)[0]

    # ---- multi-head attention with an optional relative-position window (attentions.py:231-363, vc_utils.py:568-640) ----
    # q [B,C,Tq], k / v [B,C,Tk] channel-major, heads split the channels; q_len / k_len [B] mask queries x keys (scores -1e4 where masked);
    # emb_k / emb_v [1, 2w+1, dk] or None (then plain cross / self attention)
    @staticmethod
    def _attn(q, k, v, emb_k, emb_v, q_len, k_len, heads):
        import math
        B, C, Tq = q.shape
        Tk = k.shape[2]
        dk = C // heads
        qh = q.view(B, heads, dk, Tq).transpose(2, 3) / math.sqrt(dk)
        kh = k.view(B, heads, dk, Tk).transpose(2, 3)
        vh = v.view(B, heads, dk, Tk).transpose(2, 3)
        scores = qh @ kh.transpose(-2, -1)
        if emb_k is not None:
            w = (emb_k.shape[1] - 1) // 2
            idx = torch.arange(Tk)[None, :] - torch.arange(Tq)[:, None]
            ok = (idx.abs() <= w)[:, :, None]
            rk = emb_k[0][(idx + w).clamp(0, 2 * w)] * ok
            rv = emb_v[0][(idx + w).clamp(0, 2 * w)] * ok
            scores = scores + torch.einsum("bhid,ijd->bhij", qh, rk)
        mask = (torch.arange(Tq)[None, :, None] < q_len[:, None, None]) & (torch.arange(Tk)[None, None, :] < k_len[:, None, None])
        p = torch.softmax(scores.masked_fill(~mask[:, None], -1e4), dim=-1)
        out = p @ vh
        if emb_k is not None:
            out = out + torch.einsum("bhij,ijd->bhid", p, rv)
        return out.transpose(2, 3).reshape(B, C, Tq)

    def attn_fwd(self, q, k, v, emb_k, emb_v, q_len, k_len, heads):
        return self._attn(q, k, v, emb_k, emb_v, q_len, k_len, heads)

    def attn_bwd(self, do, q, k, v, emb_k, emb_v, q_len, k_len, heads):
        if emb_k is None:
            dq, dk, dv = _vjp(lambda a, b, c: self._attn(a, b, c, None, None, q_len, k_len, heads), [q, k, v], do)
            return dq, dk, dv, None, None
        return _vjp(lambda a, b, c, d, e: self._attn(a, b, c, d, e, q_len, k_len, heads), [q, k, v, emb_k, emb_v], do)

    # ---- LayerNorm over the channel axis of [B, C, T] (modules.py:20-32), eps 1e-5 ----
    @staticmethod
    def _lnc(x, gamma, beta):
        return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, 1e-5).transpose(1, -1)

    def lnc_fwd(self, x, gamma, beta):
        return self._lnc(x, gamma, beta)

    def lnc_bwd(self, dy, x, gamma, beta):
        return _vjp(self._lnc, [x, gamma, beta], dy)

    # ---- VQ lookup with straight-through output and commitment loss (core_vq.py:174-182, 303-322): x [B,D,N], embed [K,D] ----
    @staticmethod
    def _vq_codes(x, embed):
        B, D, N = x.shape
        flat = x.permute(0, 2, 1).reshape(B * N, D)
        dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ embed.t() + embed.pow(2).sum(1)[None])
        return dist.max(dim=-1).indices

    def vq_fwd(self, x, embed):
        B, D, N = x.shape
        codes = self._vq_codes(x, embed)
        q = embed[codes].view(B, N, D).permute(0, 2, 1)
        return x + (q - x), ((q - x) ** 2).mean().reshape(1), codes

    def vq_bwd(self, dq, dcommit, x, embed, codes):
        B, D, N = x.shape
        q = embed[codes].view(B, N, D).permute(0, 2, 1)
        dx = torch.zeros_like(x)
        if dq is not None:
            dx = dx + dq                                             # straight-through
        if dcommit is not None:
            dx = dx + dcommit * 2 * (x - q) / x.numel()              # mse(q.detach(), x)
        return dx

    # ---- adversarial losses (losses.py:7-44): scalars as [1] tensors                         ttts_lsgan_loss / ttts_l1_mean ----
    def lsgan_fwd(self, x, c):
        return ((c - x) ** 2).mean().reshape(1)

    def lsgan_bwd(self, dL, x, c):
        return dL * 2 * (x - c) / x.numel()

    @staticmethod
    def _kl(z_p, logs_q, m_p, logs_p, mask):                          # mask [B, T]
        m3 = mask[:, None, :]
        kl = logs_p - logs_q - 0.5 + 0.5 * ((z_p - m_p) ** 2) * torch.exp(-2.0 * logs_p)
        return (torch.sum(kl * m3) / torch.sum(m3)).reshape(1)

    def kl_fwd(self, z_p, logs_q, m_p, logs_p, mask):
        return self._kl(z_p, logs_q, m_p, logs_p, mask)

    def kl_bwd(self, dL, z_p, logs_q, m_p, logs_p, mask):
        return _vjp(lambda a, b, c, d: self._kl(a, b, c, d, mask), [z_p, logs_q, m_p, logs_p], dL)

    def l1_fwd(self, a, b):
        return (a - b).abs().mean().reshape(1)

    def l1_bwd(self, dL, a, b):                                       # gradient of b; a is the detached (real) side
        return dL * torch.sign(b - a) / b.numel()

    # ---- transposed convolution: y = conv_transpose1d(x, w[Cin, Cout, K], b, stride, padding)   (Generator.ups, vq2.py:369-378) ----
    def convT_fwd(self, x, w, b, stride, pad):
        return F.conv_transpose1d(x, w, b, stride=stride, padding=pad)

    def convT_bwd(self, dy, x, w, stride, pad, need_db):
        dx, dw = _vjp(lambda a, c: F.conv_transpose1d(a, c, None, stride=stride, padding=pad), [x, w], dy)
        return dx, dw, (dy.sum(dim=(0, 2)) if need_db else None)

    # ---- weight norm over dim 0: w = g * v / ||v||                                           ttts_weight_norm / ttts_weight_norm_bwd ----
    @staticmethod
    def _wn(v, g):
        return g.view(-1, 1, 1) * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)

    def wn_fwd(self, v, g):
        return self._wn(v, g)

    def wn_bwd(self, dw, v, g):
        dv, dg = _vjp(self._wn, [v, g], dw)
        return dv, dg

    # ---- elementwise ----
    def add(self, a, b):
        return a + b

    def scale(self, a, s):
        return a * s

    def mul_mask(self, a, mask):                                      # a [B,C,T], mask [B,T]
        return a * mask[:, None, :]

    def lrelu_fwd(self, x, slope):
        return F.leaky_relu(x, slope)

    def lrelu_bwd(self, dy, x, slope):
        return dy * torch.where(x > 0, torch.ones_like(x), torch.full_like(x, slope))

    def tanh_fwd(self, x):
        return torch.tanh(x)

    def tanh_bwd(self, dy, x):
        return dy * (1 - torch.tanh(x) ** 2)

    def add_bcast_fwd(self, x, c):                                    # x [B,C,T] + c [B,C,1]
        return x + c

    def add_bcast_bwd(self, dy):                                      # gradient of c
        return dy.sum(-1, keepdim=True)

    @staticmethod
    def _glu(raw):
        a, g = raw.chunk(2, 1)
        return a * torch.sigmoid(g)

    def glu_fwd(self, raw):
        return self._glu(raw)

    def glu_bwd(self, dy, raw):
        return _vjp(self._glu, [raw], dy)[0]

    @staticmethod
    def _mish(x):
        return x * torch.tanh(F.softplus(x))

    def mish_fwd(self, x):
        return self._mish(x)

    def mish_bwd(self, dy, x):
        return _vjp(self._mish, [x], dy)[0]

    # ---- WN gate: tanh(a + cond_a) * sigmoid(b + cond_b), raw [B,2H,T], cond [B,2H] or None  (modules.py:195-201 fused_add_tanh_sigmoid_multiply) ----
    @staticmethod
    def _gate(raw, cond):
        a = raw if cond is None else raw + cond[:, :, None]
        h = a.shape[1] // 2
        return torch.tanh(a[:, :h]) * torch.sigmoid(a[:, h:])

    def gate_fwd(self, raw, cond):
        return self._gate(raw, cond)

    def gate_bwd(self, dy, raw, cond):
        if cond is None:
            return _vjp(lambda r: self._gate(r, None), [raw], dy)[0], None
        return _vjp(self._gate, [raw, cond], dy)

    # ---- Activation1d(SnakeBeta, log-scale alpha / beta)                                      alias_free_torch/act.py:8-28, activations.py:62-119 ----
    @staticmethod
    def _snake(x, la, lb, filt):
        C = x.shape[1]
        f = filt.view(1, 1, -1)
        xp = F.pad(x, (5, 5), mode="replicate")
        up = 2 * F.conv_transpose1d(xp, f.expand(C, -1, -1), stride=2, groups=C)[..., 15:-15]
        alpha, beta = torch.exp(la).view(1, -1, 1), torch.exp(lb).view(1, -1, 1)
        up = up + (1.0 / (beta + 1e-9)) * torch.sin(up * alpha) ** 2
        return F.conv1d(F.pad(up, (5, 6), mode="replicate"), f.expand(C, -1, -1), stride=2, groups=C)

    def snake_fwd(self, x, la, lb, filt):
        return self._snake(x, la, lb, filt)

    def snake_bwd(self, dy, x, la, lb, filt):
        return _vjp(lambda a, b, c: self._snake(a, b, c, filt), [x, la, lb], dy)

    # ---- small masked multi-head attention, channel-major [B,C,T]; keys >= lens[b] masked    modules.py:640-683 (MultiHeadAttention of MelStyleEncoder) ----
    @staticmethod
    def _mha(q, k, v, lens, heads, temperature):
        B, C, T = q.shape
        dk = C // heads
        qh, kh, vh = [t.view(B, heads, dk, T) for t in (q, k, v)]
        att = torch.einsum("bhdt,bhds->bhts", qh, kh) / temperature
        pad = torch.arange(T)[None, :] >= lens[:, None]                     # [B, T] True at padding
        att = att.masked_fill(pad[:, None, None, :], float("-inf"))
        return torch.einsum("bhts,bhds->bhdt", torch.softmax(att, dim=-1), vh).reshape(B, C, T)

    def mha_fwd(self, q, k, v, lens, heads, temperature):
        return self._mha(q, k, v, lens, heads, temperature)

    def mha_bwd(self, do, q, k, v, lens, heads, temperature):
        return _vjp(lambda a, b, c: self._mha(a, b, c, lens, heads, temperature), [q, k, v], do)

    # ---- masked temporal mean: y[b,c] = sum_{t < lens[b]} x[b,c,t] / lens[b]                  modules.py:757-763 ----
    def masked_mean_fwd(self, x, lens):
        T = x.shape[-1]
        m = (torch.arange(T)[None, :] < lens[:, None]).float()
        return (x * m[:, None, :]).sum(-1) / lens[:, None].float()

    def masked_mean_bwd(self, dy, lens, T):
        m = (torch.arange(T)[None, :] < lens[:, None]).float()
        return (dy / lens[:, None].float())[:, :, None] * m[:, None, :]

    # ---- posterior sample: z = (m + eps * exp(logs)) * mask, stats = [m | logs]               vq2.py:742-744 ----
    @staticmethod
    def _posterior(stats, eps, mask):
        m, logs = stats.chunk(2, 1)
        e = eps if eps is not None else torch.zeros_like(m)
        return (m + e * torch.exp(logs)) * mask[:, None, :]

    def posterior_fwd(self, stats, eps, mask):
        return self._posterior(stats, eps, mask)

    def posterior_bwd(self, dz, stats, eps, mask):
        return _vjp(lambda s: self._posterior(s, eps, mask), [stats], dz)[0]

    # ================= diffusion mel-refiner (ttts_b200/diffusion/train_graph.py; csrc/diffusion_kernels.cu) =================
    # ---- act(GroupNorm(x) * (1 + scale) + shift), eps 1e-5 (utils.py:119-137, aa_model.py:126-130); stats [B, G, 2] = (mean, rstd) ----
    @staticmethod
    def _gn(x, gamma, beta, groups, scale, shift, silu):
        y = F.group_norm(x, groups, gamma, beta, eps=1e-5)
        if scale is not None:
            y = y * (1 + scale) + shift
        return F.silu(y) if silu else y

    def gn_fwd(self, x, gamma, beta, groups, scale, shift, silu):
        B, C, T = x.shape
        xg = x.reshape(B, groups, -1)
        mean = xg.mean(-1)
        rstd = torch.rsqrt(xg.var(-1, unbiased=False) + 1e-5)
        return self._gn(x, gamma, beta, groups, scale, shift, silu), torch.stack([mean, rstd], dim=-1)

    def gn_bwd(self, dy, x, stats, gamma, beta, groups, scale, shift, silu):
        if scale is None:
            dx, dg, db = _vjp(lambda a, g, b: self._gn(a, g, b, groups, None, None, silu), [x, gamma, beta], dy)
            return dx, dg, db, None, None
        return _vjp(lambda a, g, b, sc, sh: self._gn(a, g, b, groups, sc, sh, silu), [x, gamma, beta, scale, shift], dy)

    def silu_fwd(self, x):
        return F.silu(x)

    def silu_bwd(self, dy, x):
        return _vjp(F.silu, [x], dy)[0]

    # ---- QKVAttentionLegacy + RelativePositionBias (utils.py:148-175, xtransformers.py:178-188): qkv [B, 3C, T] with the heads split BEFORE
    # q / k / v, table [32, H], diag int32 [2T-1] = bucket of (j - i) + T - 1; returns (out [B,C,T], lse [B,H,T] of the biased scores) ----
    @staticmethod
    def _attn_bias(qkv, table, heads, diag, want_lse=False):
        import math
        B, W, T = qkv.shape
        ch = W // (3 * heads)
        q, k, v = qkv.reshape(B * heads, ch * 3, T).split(ch, dim=1)
        w = torch.einsum("bct,bcs->bts", q, k) / math.sqrt(ch)
        idx = (torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1)
        bias = table[diag.long()[idx]]                               # [T, T, H]
        w = (w.reshape(B, heads, T, T) + bias.permute(2, 0, 1)[None] * math.sqrt(ch)).reshape(B * heads, T, T)
        out = torch.einsum("bts,bcs->bct", torch.softmax(w, dim=-1), v).reshape(B, -1, T)
        return (out, torch.logsumexp(w, dim=-1).reshape(B, heads, T)) if want_lse else out

    def attn_bias_fwd(self, qkv, table, heads, diag):
        return self._attn_bias(qkv, table, heads, diag, True)

    def attn_bias_bwd(self, do, qkv, out, lse, table, heads, diag):
        return _vjp(lambda a, t: self._attn_bias(a, t, heads, diag), [qkv, table], do)

    # ---- q_sample and the loss (utils/diffusion.py:243-260, 903-1014); coef [B, 8] fp32, see train_graph.coef_table ----
    def q_sample(self, x_start, noise, coef):
        return coef[:, 0, None, None] * x_start + coef[:, 1, None, None] * noise

    @staticmethod
    def _diff_terms(out, x_start, x_t, noise, coef, t_is0):
        import math
        Cn = x_start.shape[1]
        co = lambda i: coef[:, i, None, None]
        eps, v = out[:, :Cn], out[:, Cn:]
        mse = ((noise - eps) ** 2).mean(dim=(1, 2))
        eps_d = eps.detach()
        true_mean = co(4) * x_start + co(5) * x_t
        frac = (v + 1) / 2
        logvar = frac * co(7) + (1 - frac) * co(6)
        mean = co(4) * (co(2) * x_t - co(3) * eps_d).clamp(-1, 1) + co(5) * x_t
        kl = 0.5 * (-1.0 + logvar - co(6) + torch.exp(co(6) - logvar) + ((true_mean - mean) ** 2) * torch.exp(-logvar))
        kl = kl.mean(dim=(1, 2)) / math.log(2.0)
        cdf = lambda a: 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (a + 0.044715 * torch.pow(a, 3))))
        cx = x_start - mean
        inv_std = torch.exp(-0.5 * logvar)
        cp, cm = cdf(inv_std * (cx + 1.0 / 255.0)), cdf(inv_std * (cx - 1.0 / 255.0))
        lp = torch.where(x_start < -0.999, torch.log(cp.clamp(min=1e-12)),
                         torch.where(x_start > 0.999, torch.log((1.0 - cm).clamp(min=1e-12)), torch.log((cp - cm).clamp(min=1e-12))))
        nll = -lp.mean(dim=(1, 2)) / math.log(2.0)
        return mse, torch.where(t_is0 != 0, nll, kl)

    def diff_loss_fwd(self, out, x_start, x_t, noise, coef, t_is0):
        mse, vb = self._diff_terms(out, x_start, x_t, noise, coef, t_is0)
        return (mse + vb).mean().reshape(1), (mse, vb)

    def diff_loss_bwd(self, dL, out, x_start, x_t, noise, coef, t_is0):
        def f(o):
            mse, vb = self._diff_terms(o, x_start, x_t, noise, coef, t_is0)
            return (mse + vb).mean().reshape(1)
        return _vjp(f, [out], dL)[0]


class TorchAdamW:
    """the optimizer contract of train_step.TrainStep with torch.optim.AdamW(lr, betas (0.8, 0.99), eps 1e-9) -- the trainer's own
    (ttts/vqvae/train.py:193-205) -- so that the ORDER of the step can be checked against the reference on CPU"""

    def __init__(self, params, lr=1e-4):
        self.p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.opt = torch.optim.AdamW(list(self.p.values()), lr, betas=(0.8, 0.99), eps=1e-9)

    def params(self):
        return {k: v.detach() for k, v in self.p.items()}              # same storage: the in-place update is visible to the graphs

    def step(self, grads):
        for k, v in self.p.items():
            v.grad = grads[k].reshape(v.shape).clone()
        self.opt.step()


class TorchAdamWClip:
    """the optimizer contract of ttts_b200.diffusion.train_step.DiffusionStep with torch's own pieces -- clip_grad_norm_(1.0) +
    torch.optim.AdamW(lr, betas (0.9, 0.999), weight decay 0.01) with the learning rate scaled per step (ttts/diffusion/train.py:116-117,190-195)"""

    def __init__(self, params, lr=1e-4):
        self.p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        self.lr = lr
        self.opt = torch.optim.AdamW(list(self.p.values()), lr, betas=(0.9, 0.999), weight_decay=0.01)

    def params(self):
        return {k: v.detach() for k, v in self.p.items()}

    def step(self, grads, lr_scale=1.0):
        for k, v in self.p.items():
            v.grad = grads[k].reshape(v.shape).clone()
        norm = torch.nn.utils.clip_grad_norm_(list(self.p.values()), 1.0)
        for g in self.opt.param_groups:
            g["lr"] = self.lr * lr_scale
        self.opt.step()
        return norm.reshape(1)
