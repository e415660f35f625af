"""`AA_diffusion.forward` in train() mode and the diffusion loss on the training tape (ttts_b200/vqvae/train_encoder.py: `Tape`, `Ops`):
every op is ONE forward kernel call on the backend `K` and ONE recorded closure calling its backward kernel; no torch.autograd on the compute
path.  Reference: ttts/diffusion/aa_model.py:69-287 (ResBlock with scale-shift norm, DiffusionLayer, RefEncoder, AA_diffusion),
ttts/utils/utils.py:119-215 (GroupNorm32, QKVAttentionLegacy, AttentionBlock), utils/xtransformers.py:146-188 (RelativePositionBias),
utils/vc_utils.py:514-600 (the latents' cross attention), utils/diffusion.py:903-1014 (loss).  Ops this graph adds to the contract:
`gn` (GroupNorm + optional (1 + scale) / shift modulation + optional SiLU, fused), `silu`, `attn_bias` (non-causal attention on the packed
head-major qkv with the bucketed relative-position bias), `diff_loss` (MSE + variational-bound term), `q_sample`.  The model's random
decisions (unconditioned samples, dropped layers, t, noise) are inputs -- `DiffusionStep` draws them.

tests/ref_kernels.py restates the contract of the new ops in torch; tests/test_train_diffusion_cpu.py runs THIS graph over it against the REAL
reference's micro-step (tests/golden/diffusion.npz), which pins the wiring; the CUDA kernels (csrc/diffusion_kernels.cu) are checked op by op
against the same contract on the CPU emulation and on the GPU."""
import math

import numpy as np
import torch

from ..vqvae.train_encoder import Ops, Tape, Var

N_LATENTS, REF_HEADS = 32, 8
NUM_BUCKETS, MAX_DISTANCE = 32, 64


def gn_groups(channels):
    """`normalization` (ttts/utils/utils.py:124-137)"""
    groups = 32
    if channels <= 16:
        groups = 8
    elif channels <= 64:
        groups = 16
    while channels % groups != 0:
        groups = int(groups / 2)
    assert groups > 2
    return groups


def diagonal_buckets(T):
    """bucket of the relative position j - i for j - i = -(T-1) .. T-1 (xtransformers.py:155-176, causal=False, 32 buckets, max_distance 64):
    an int32 [2T-1] table indexed by (j - i) + T - 1.  Host-side integer constant of the shape, computed once per length."""
    rel = np.arange(-(T - 1), T)
    n = -rel
    nb = NUM_BUCKETS // 2
    ret = (n < 0).astype(np.int64) * nb
    n = np.abs(n)
    max_exact = nb // 2
    with np.errstate(divide="ignore"):
        large = max_exact + (np.log(np.maximum(n, 1).astype(np.float32) / np.float32(max_exact)) / np.float32(math.log(MAX_DISTANCE / max_exact)) * (nb - max_exact)).astype(np.int64)
    large = np.minimum(large, nb - 1)
    return torch.tensor((ret + np.where(n < max_exact, n, large)).astype(np.int32))


def nearest_index(T_in, T_out):
    """source index of F.interpolate(mode="nearest") (aa_model.py:253): floor(dst * T_in / T_out) in fp32 like ATen"""
    scale = np.float32(T_in) / np.float32(T_out)
    return torch.tensor(np.minimum(np.floor(np.arange(T_out, dtype=np.float32) * scale).astype(np.int64), T_in - 1))


def timestep_embedding(t, dim, max_period=10000):
    """aa_model.py:33-52 (a constant of the step: no parameters)"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t.cpu()[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def schedule(n=1000):
    """linear betas and the tables GaussianDiffusion.__init__ derives (utils/diffusion.py:92-98, 202-229), float64 like the reference's"""
    betas = np.linspace(1000 / n * 0.0001, 1000 / n * 0.02, n, dtype=np.float64)
    ac = np.cumprod(1.0 - betas)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return np.stack([np.sqrt(ac), np.sqrt(1.0 - ac), np.sqrt(1.0 / ac), np.sqrt(1.0 / ac - 1), betas * np.sqrt(ac_prev) / (1.0 - ac),
                     (1.0 - ac_prev) * np.sqrt(1.0 - betas) / (1.0 - ac), np.log(np.append(post_var[1], post_var[1:])), np.log(betas)], axis=1)


_SCHEDULE = None


def coef_table(t):
    """fp32 [B, 8] per-sample coefficients (`_extract_into_tensor(...).float()`): sqrt_ac, sqrt_1mac, sqrt_recip_ac, sqrt_recipm1_ac,
    posterior_mean_coef1, coef2, min_log = posterior_log_variance_clipped, max_log = log beta"""
    global _SCHEDULE
    if _SCHEDULE is None:
        _SCHEDULE = schedule()
    return torch.tensor(_SCHEDULE[np.asarray(t.cpu() if torch.is_tensor(t) else t)].astype(np.float32))


class DiffOps(Ops):
    def gn(self, x, gamma, beta, groups, scale=None, shift=None, silu=False):
        """act(GroupNorm(x) * (1 + scale) + shift): x [B,C,T], scale / shift [B,C,1] Vars or None, act = SiLU or identity"""
        sv, hv = (scale.v, shift.v) if scale is not None else (None, None)
        yv, stats = self.K.gn_fwd(x.v, gamma.v, beta.v, groups, sv, hv, silu)
        y = Var(yv)

        def bwd():
            if y.g is None:
                return
            dx, dg, db, dsc, dsh = self.K.gn_bwd(y.g, x.v, stats, gamma.v, beta.v, groups, sv, hv, silu)
            self._acc(x, dx); self._acc(gamma, dg); self._acc(beta, db)
            if scale is not None:
                self._acc(scale, dsc); self._acc(shift, dsh)
        self.tape.record(bwd)
        return y

    def silu(self, x):
        return self._unary("silu", x)

    def attn_bias(self, qkv, table, heads, diag):
        """QKVAttentionLegacy + RelativePositionBias: qkv [B, 3C, T] (per head: q | k | v blocks of ch channels), table [32, H] Var,
        diag = diagonal_buckets(T) on the device -> [B, C, T]"""
        ov, lse = self.K.attn_bias_fwd(qkv.v, table.v, heads, diag)
        o = Var(ov)

        def bwd():
            if o.g is None:
                return
            dqkv, dtab = self.K.attn_bias_bwd(o.g, qkv.v, ov, lse, table.v, heads, diag)
            self._acc(qkv, dqkv); self._acc(table, dtab)
        self.tape.record(bwd)
        return o

    def diff_loss(self, out, x_start, x_t, noise, coef, t_is0):
        """mean over the batch of mse + vb (utils/diffusion.py:975-1010, train.py:172-180) as a [1] Var; also returns the per-sample terms"""
        lv, terms = self.K.diff_loss_fwd(out.v, x_start, x_t, noise, coef, t_is0)
        y = Var(lv)

        def bwd():
            if y.g is not None:
                self._acc(out, self.K.diff_loss_bwd(y.g, out.v, x_start, x_t, noise, coef, t_is0))
        self.tape.record(bwd)
        return y, terms

    # ---- memory plumbing ----
    def gather_t(self, x, idx):
        """x[:, :, idx] (nearest-neighbour interpolation); the backward scatter-adds"""
        y = Var(x.v.index_select(2, idx).contiguous())

        Tin, Tout = x.v.shape[2], idx.numel()

        def bwd():
            if y.g is None:
                return
            if Tout % Tin == 0:                                      # integer ratio: every source frame is repeated Tout / Tin times in a row
                r = Tout // Tin
                g = y.g[:, :, 0::r].contiguous()
                for k in range(1, r):
                    g = self.K.add(g, y.g[:, :, k::r].contiguous())   # fixed order (index_add_ on the GPU is atomic)
            else:
                g = torch.zeros_like(x.v)
                g.index_add_(2, idx, y.g)
            self._acc(x, g)
        self.tape.record(bwd)
        return y

    def cat_t(self, a, b):
        ta = a.v.shape[2]
        y = Var(torch.cat([a.v, b.v], dim=2))

        def bwd():
            if y.g is not None:
                self._acc(a, y.g[:, :, :ta].contiguous()); self._acc(b, y.g[:, :, ta:].contiguous())
        self.tape.record(bwd)
        return y

    def select_batch(self, x, alt, mask):
        """torch.where(mask[b], alt, x[b]) (aa_model.py:250-252): alt [1,C,1] Var broadcast over the masked samples and time; mask: CPU bool [B]"""
        B, C, T = x.v.shape
        dev = x.v.device
        masked = [b for b in range(B) if bool(mask[b])]
        keep = torch.tensor([0.0 if bool(mask[b]) else 1.0 for b in range(B)], device=dev)[:, None].expand(B, T).contiguous()
        y = Var(torch.where(mask.to(dev)[:, None, None], alt.v.expand(B, C, T), x.v).contiguous())

        def bwd():
            if y.g is None:
                return
            self._acc(x, self.K.mul_mask(y.g, keep))
            if masked:
                s = self.K.add_bcast_bwd(y.g)                      # [B,C,1] sums over time
                acc = None
                for b in masked:
                    acc = s[b:b + 1].contiguous() if acc is None else self.K.add(acc, s[b:b + 1].contiguous())
                self._acc(alt, acc)
        self.tape.record(bwd)
        return y

    def expand_latents(self, lat, B):
        """[n, d] parameter -> [B, d, n] (einops repeat "n d -> b d n", aa_model.py:170)"""
        y = Var(lat.v.t()[None].expand(B, -1, -1).contiguous())

        def bwd():
            if y.g is not None:
                acc = y.g[0]
                for b in range(1, B):
                    acc = self.K.add(acc, y.g[b])
                self._acc(lat, acc.t().contiguous())
        self.tape.record(bwd)
        return y


class DiffusionGraph:
    """AA_diffusion over the reference's state_dict names.  `cfg`: model_channels, num_layers, num_heads (aa_model.py:183-195)."""

    def __init__(self, K, params, cfg, tape=None):
        self.K, self.cfg = K, cfg
        self.tape = tape if tape is not None else Tape()
        self.ops = DiffOps(K, self.tape)
        self.P = {k: Var(v.detach().unsqueeze(-1).contiguous() if (v.dim() == 2 and not k.endswith("latents") and "relative_attention_bias" not in k)
                         else v.detach().contiguous()) for k, v in params.items()}
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}
        self._diag = {}

    def diag(self, T, dev):
        if T not in self._diag:
            self._diag[T] = diagonal_buckets(T).to(dev)
        return self._diag[T]

    def conv(self, pre, x, pad=0, need_dx=True):
        return self.ops.conv(x, self.P[pre + "weight"], self.P[pre + "bias"], pad=pad, need_dx=need_dx)

    def gn(self, pre, x, scale=None, shift=None, silu=False):
        return self.ops.gn(x, self.P[pre + "weight"], self.P[pre + "bias"], gn_groups(x.v.shape[1]), scale, shift, silu)

    def attention_block(self, pre, x, heads):
        """utils.py:209-215"""
        o = self.ops
        qkv = self.conv(pre + "qkv.", self.gn(pre + "norm.", x))
        h = o.attn_bias(qkv, self.P[pre + "relative_pos_embeddings.relative_attention_bias.weight"], heads, self.diag(x.v.shape[2], x.v.device))
        return o.add(x, self.conv(pre + "proj_out.", h))

    def resblock(self, pre, x, emb_act):
        """aa_model.py:120-135; emb_act = SiLU(time embedding) [B,C,1] (shared by every block: emb_layers.0 is a parameter-free SiLU)"""
        o = self.ops
        C = x.v.shape[1]
        h = self.conv(pre + "in_layers.2.", self.gn(pre + "in_layers.0.", x, silu=True))
        eo = self.conv(pre + "emb_layers.1.", emb_act)
        h = self.gn(pre + "out_layers.0.", h, o.slice_c(eo, 0, C), o.slice_c(eo, C, 2 * C), silu=True)
        return o.add(x, self.conv(pre + "out_layers.3.", h, pad=1))

    def diffusion_layer(self, pre, x, emb_act, heads):
        return self.attention_block(pre + "attn.", self.resblock(pre + "resblk.", x, emb_act), heads)

    def ref_encoder(self, pre, x):
        """aa_model.py:153-177"""
        o = self.ops
        B, C, T = x.v.shape
        lat = o.expand_latents(self.P[pre + "latents"], B)
        q = self.conv(pre + "cross_attention.conv_q.", lat)
        k = self.conv(pre + "cross_attention.conv_k.", x)
        v = self.conv(pre + "cross_attention.conv_v.", x)
        dev = x.v.device
        ql = torch.full((B,), N_LATENTS, dtype=torch.int64, device=dev)
        kl = torch.full((B,), T, dtype=torch.int64, device=dev)
        lat = self.conv(pre + "cross_attention.conv_o.", o.attn(q, k, v, None, None, ql, kl, REF_HEADS))
        h = self.conv(pre + "enc.0.", o.cat_t(lat, x), pad=1)
        for i in (1, 2, 3, 4):
            h = self.attention_block(pre + "enc.%d." % i, h, REF_HEADS)
        Tt = h.v.shape[2]
        return o.reshape(o.masked_mean(h, torch.full((B,), Tt, dtype=torch.int64, device=dev)), (B, C, 1))

    def forward(self, x_t, t, latent, refer, uncond=None, dropped=()):
        """aa_model.py:256-287 (train mode).  x_t [B,100,T], t [B] int64, latent [B,Cl,Tl], refer [B,100,Tr] (constants); uncond bool [B] or
        None; dropped: indices of `layers` skipped this step.  Returns the Var of the model output [B,200,T]."""
        o, cfg = self.ops, self.cfg
        C, H, L = cfg["model_channels"], cfg["num_heads"], cfg["num_layers"]
        dev = x_t.device
        B, _, T = x_t.shape
        h = self.conv("latent_conditioner.0.", Var(latent.contiguous()), pad=1, need_dx=False)
        for i in (1, 2, 3):
            h = self.attention_block("latent_conditioner.%d." % i, h, H)
        r = self.conv("refer_enc.0.", Var(refer.contiguous()), pad=1, need_dx=False)
        for i in (1, 2, 3):
            r = self.attention_block("refer_enc.%d." % i, r, H)
        r = self.ref_encoder("refer_enc.4.", r)
        le = o.add_bcast(self.gn("code_norm.", h), r)
        if uncond is not None:
            le = o.select_batch(le, self.P["unconditioned_embedding"], uncond.cpu())
        le = o.gather_t(le, nearest_index(le.v.shape[2], T).to(dev))
        te0 = Var(timestep_embedding(t, C).to(dev).unsqueeze(-1).contiguous())
        te = self.conv("time_embed.2.", o.silu(self.conv("time_embed.0.", te0, need_dx=False)))
        emb_act = o.silu(te)
        for i in range(3):
            le = self.diffusion_layer("conditioning_timestep_integrator.%d." % i, le, emb_act, H)
        x = self.conv("inp_block.", Var(x_t.contiguous()), pad=1, need_dx=False)
        x = self.conv("integrating_conv.", o.cat_c(x, le))
        for i in range(L + 3):
            if i in dropped:
                assert 0 < i < L + 2, "the first and the last layer are never dropped (aa_model.py:270)"
                continue
            x = self.diffusion_layer("layers.%d." % i, x, emb_act, H) if i < L else self.resblock("layers.%d." % i, x, emb_act)
        return self.conv("out.2.", self.gn("out.0.", x, silu=True), pad=1)

    def loss(self, x_start, t, noise, latent, refer, uncond=None, dropped=()):
        """one micro-step of ttts/diffusion/train.py:168-180: q_sample -> model -> training_losses(...)["loss"].mean().
        Returns (loss Var [1], dict(mse [B], vb [B], model_out Var))."""
        dev = x_start.device
        coef = coef_table(t).to(dev)
        t_is0 = (t == 0).to(device=dev, dtype=torch.int32).contiguous()
        x_t = self.K.q_sample(x_start.contiguous(), noise.contiguous(), coef)
        out = self.forward(x_t, t, latent, refer, uncond, dropped)
        lossv, terms = self.ops.diff_loss(out, x_start.contiguous(), x_t, noise.contiguous(), coef, t_is0)
        return lossv, dict(mse=terms[0], vb=terms[1], model_out=out)

    def backward(self, lossv):
        lossv.g = torch.ones_like(lossv.v)
        self.tape.backward()
        return self.grads()

    def grads(self):
        """name -> gradient in the reference's state_dict shape; parameters untouched by this step (dropped layers, the unconditioned embedding
        when no sample drew it) get zeros, like the reference's `extraneous_addition * 0` (aa_model.py:281-285)"""
        return {k: (v.g.reshape(self.shapes[k]) if v.g is not None else torch.zeros(self.shapes[k], device=v.v.device)) for k, v in self.P.items()}
