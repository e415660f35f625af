"""VQ-VAE encode front end (forward / extraction path): `MelStyleEncoder`, `WN`, `ResBlock1`, `PosteriorAudioEncoder` and the
`VQEncoder` pipeline = the first half of `SynthesizerTrn.forward / infer` (ttts/vqvae/vq2.py:843-852, 874-882):

    spec = spectrogram_torch(wav) ; ge = ref_enc(spec*mask, mask) ; x,_,_ = enc_p(spec, wav, mask, g=ge) ; x = proj(x) ;
    quantized, codes, commit, _ = quantizer(x, layers=[0])

Module / parameter names follow the reference so that `load_state_dict` accepts a reference checkpoint's `ref_enc.*`, `enc_p.*`,
`proj.*`, `quantizer.*` entries unchanged -- including BOTH weight-norm spellings the reference mixes:
old-style `weight_g / weight_v` (`WN`, `downs`) and `parametrizations.weight.original0 / original1` (`ResBlock1`).

All arithmetic runs in csrc/conv1d.cu (fp32 direct convolutions with the blocks' elementwise work fused), csrc/stft.cu and
csrc/vq.cu.  Forward only (`torch.no_grad()` semantics): the encoder's backward belongs to the VQ-VAE-GAN train step, a "next"
row of SURVEY.md section 8(f).  No CPU fallback.
"""
import ctypes
import math
import weakref
import os

import torch
from torch import nn

from .. import _lib as L
from .mel import spectrogram_torch
from .quantize import ResidualVectorQuantizer


def _protos(lib):
    if getattr(lib, "_conv_protos", False):
        return
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
    lib.ttts_conv1d_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, f32, i32, vp, i32, vp, i32, vp]
    lib.ttts_conv1d_f32_split.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, f32, i32, vp, i32, vp, i32, i32, vp]
    lib.ttts_conv1d_tcs_weight_elems.argtypes = [i32, i32, i32]
    lib.ttts_conv1d_tcs_weight_elems.restype = ctypes.c_int64
    lib.ttts_conv1d_tcs_prep_weights.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.ttts_conv1d_tcs.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, f32, i32, vp, i32, vp, i32, i32, vp]
    lib.ttts_conv1d_bwd_input.argtypes = [vp, vp, vp, vp] + [i32] * 10 + [vp]
    lib.ttts_conv1d_bwd_weight.argtypes = [vp, vp, vp, vp] + [i32] * 9 + [vp]
    lib.ttts_weight_norm.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.ttts_snake_aa.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.ttts_mha_small.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]
    lib.ttts_masked_mean.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.ttts_posterior_sample.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib._conv_protos = True


def _p(t):
    return t.data_ptr() if t is not None else None


# True while a VQEncoder forward with `conv_tc` set runs: every convolution the tensor-core kernel covers (csrc/conv1d_tcs.cu) goes to it.
# Everything else that calls conv1d() -- the training tape (ttts_b200/vqvae/train_*.py), the per-kernel tests -- keeps the exact-fp32 kernels.
USE_TC = False
TC_FLAGS = int(os.environ.get("TTTS_CONV_TC_FLAGS", "0"))


def tcs_covers(Cin, Cout, K, stride, dil, pad, post, cond):
    if not (stride == 1 and 2 * pad == dil * (K - 1) and 128 + 2 * pad <= 184 and Cin % 8 == 0 and 16 <= Cin <= 192):
        return False
    if post == 3:
        return Cout == 384                                    # the WN gate: 192 tanh channels x 192 sigmoid channels
    return post == 0 and cond is None and Cout in (32, 64, 96, 128, 192, 384)


def tcs_weights(w):
    """The split-bf16 form [tap][ci atom][hi | lo][co][64] of a convolution weight for ttts_conv1d_tcs, cached ON the weight tensor object
    (validated by address + version + shape, as weight_norm_apply does): constant weights are split once, not once per call."""
    state = (w.data_ptr(), w._version, tuple(w.shape))
    capturing = torch.cuda.is_current_stream_capturing()
    hit = getattr(w, "_ttts_tcs", None)
    if hit is not None and hit[0] == state:                # also while a CUDA graph is being captured: the cached tensor predates the capture
        return hit[1]
    lib = L.lib(); _protos(lib)
    Cout, Cin, K = w.shape
    ws = torch.empty(lib.ttts_conv1d_tcs_weight_elems(Cout, Cin, K), dtype=torch.bfloat16, device=w.device)
    L.check(lib.ttts_conv1d_tcs_prep_weights(_p(w), _p(ws), Cout, Cin, K, L.stream_ptr().value), "ttts_conv1d_tcs_prep_weights")
    if not capturing:
        w._ttts_tcs = (state, ws)
    return ws


def conv1d(x, w, bias=None, stride=1, dil=1, pad=0, pre_lrelu=False, resid=None, out_scale=1.0, out=None, accumulate=False, mask=None,
           post=0, cond=None, split=0, tc=None):
    """Raw call of ttts_conv1d_f32.  x [B,Cin,T] fp32 contiguous, w [Cout,Cin,K].  split = 2 / 4: force the split-reduction kernel
    (ttts_conv1d_f32_split; per-kernel tests); tc: True = the split-bf16 tcgen05 kernel (ttts_conv1d_tcs), False = the fp32 kernels,
    None = the tensor-core kernel when TTTS_CONV_TC=1 and it covers the layer."""
    lib = L.lib(); _protos(lib)
    L.require_cuda(x, w)
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.float32 and w.dtype == torch.float32
    B, Cin, Tin = x.shape
    Cout, Cin2, K = w.shape
    assert Cin == Cin2
    Tout = (Tin + 2 * pad - dil * (K - 1) - 1) // stride + 1
    Ceff = Cout // 2 if post in (1, 3) else Cout
    if out is None:
        out = torch.empty(B, Ceff, Tout, dtype=torch.float32, device=x.device)
    cond_ld = cond.stride(0) if cond is not None else 0
    if tc is None:
        tc = USE_TC and not split and tcs_covers(Cin, Cout, K, stride, dil, pad, post, cond)
    if tc:
        assert tcs_covers(Cin, Cout, K, stride, dil, pad, post, cond), "layer not covered by ttts_conv1d_tcs"
        L.check(lib.ttts_conv1d_tcs(_p(x), _p(tcs_weights(w)), _p(bias), _p(out), B, Cin, Tin, Cout, K, dil, int(pre_lrelu), _p(resid), float(out_scale),
                                    int(accumulate), _p(mask), post, _p(cond), cond_ld, TC_FLAGS, L.stream_ptr().value), "ttts_conv1d_tcs")
        return out
    if split:
        L.check(lib.ttts_conv1d_f32_split(_p(x), _p(w), _p(bias), _p(out), B, Cin, Tin, Cout, K, stride, dil, pad, int(pre_lrelu), _p(resid),
                                          float(out_scale), int(accumulate), _p(mask), post, _p(cond), cond_ld, int(split), L.stream_ptr().value),
                "ttts_conv1d_f32_split")
        return out
    L.check(lib.ttts_conv1d_f32(_p(x), _p(w), _p(bias), _p(out), B, Cin, Tin, Cout, K, stride, dil, pad, int(pre_lrelu), _p(resid), float(out_scale),
                                int(accumulate), _p(mask), post, _p(cond), cond_ld, L.stream_ptr().value), "ttts_conv1d_f32")
    return out


def conv1d_backward(dy, x, w, stride=1, dil=1, pad=0, pre_lrelu=False, need_bias=True):
    """Autograd of `conv1d(x, w, b, stride, dil, pad, pre_lrelu)` for a given dy [B,Cout,Tout]: returns (dx, dw, db).  Raw calls of
    ttts_conv1d_bwd_input / ttts_conv1d_bwd_weight (csrc/conv1d_bwd.cu; used by the training tape, ttts_b200/vqvae/train_encoder.py)."""
    lib = L.lib(); _protos(lib)
    L.require_cuda(dy, x, w)
    assert dy.is_contiguous() and x.is_contiguous() and w.is_contiguous()
    B, Cin, Tin = x.shape
    Cout, _, K = w.shape
    dx = torch.empty_like(x)
    dw = torch.zeros_like(w)
    db = torch.zeros(Cout, dtype=torch.float32, device=x.device) if need_bias else None
    st = L.stream_ptr().value
    L.check(lib.ttts_conv1d_bwd_input(_p(dy), _p(w), _p(x), _p(dx), B, Cin, Tin, Cout, K, stride, dil, pad, int(pre_lrelu), 0, st), "ttts_conv1d_bwd_input")
    L.check(lib.ttts_conv1d_bwd_weight(_p(dy), _p(x), _p(dw), _p(db), B, Cin, Tin, Cout, K, stride, dil, pad, int(pre_lrelu), st), "ttts_conv1d_bwd_weight")
    return dx, dw, db


def weight_norm_apply(v, g):
    """g * v / ||v|| (per output channel).  Cached ON the parameter object `v` (attribute `_ttts_wn`), validated by the identity of `g`,
    storage addresses, version counters and shape (so `.cuda()`, `load_state_dict`, an optimizer step or any in-place edit invalidate it):
    in eval / extraction the weights are constants, so the ~170 normalisations of the encoder stack run once instead of once per call.
    Living on the object (not in a table keyed by data_ptr()) the entry dies with its module, and a second model that the caching
    allocator places at a freed model's addresses can never hit it."""
    state = (v.data_ptr(), g.data_ptr(), v._version, g._version, tuple(v.shape))
    capturing = torch.cuda.is_current_stream_capturing()
    hit = getattr(v, "_ttts_wn", None)
    if hit is not None and hit[0]() is g and hit[1] == state:      # also under graph capture (entries are only WRITTEN outside a capture)
        return hit[2]
    lib = L.lib(); _protos(lib)
    w = torch.empty_like(v)
    L.check(lib.ttts_weight_norm(_p(v), _p(g), _p(w), v.shape[0], v[0].numel(), L.stream_ptr().value), "ttts_weight_norm")
    if not capturing:
        v._ttts_wn = (weakref.ref(g), state, w)
    return w


def kaiser_sinc_filter12():
    """alias_free_torch/filter.py:29-58 with cutoff 0.25, half_width 0.3, kernel_size 12."""
    cutoff, half_width, ks = 0.25, 0.3, 12
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    window = torch.kaiser_window(ks, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (f / f.sum()).float()


# ------------------------------------------------------------------------------------------------ parameter holders
class _ConvWN(nn.Module):
    """Conv1d under OLD torch.nn.utils.weight_norm: parameters weight_g [Cout,1,1], weight_v [Cout,Cin,K], bias."""

    def __init__(self, cin, cout, k, stride=1, dil=1, pad=0):
        super().__init__()
        self.stride, self.dil, self.pad = stride, dil, pad
        v = torch.empty(cout, cin, k)
        nn.init.kaiming_uniform_(v, a=math.sqrt(5))
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(cout, 1, 1))
        self.weight_v = nn.Parameter(v)
        bound = 1.0 / math.sqrt(cin * k)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))

    def weight(self):
        return weight_norm_apply(self.weight_v, self.weight_g)


class _Orig(nn.Module):
    def __init__(self, g, v):
        super().__init__()
        self.original0 = nn.Parameter(g)
        self.original1 = nn.Parameter(v)


class _ConvPWN(nn.Module):
    """Conv1d under torch.nn.utils.parametrizations.weight_norm: parametrizations.weight.original0 (g), original1 (v), bias."""

    def __init__(self, cin, cout, k, dil=1, pad=0):
        super().__init__()
        self.dil, self.pad = dil, pad
        v = torch.empty(cout, cin, k).normal_(0.0, 0.01)          # commons.init_weights (mean 0, std 0.01)
        self.parametrizations = nn.Module()
        self.parametrizations.weight = _Orig(v.flatten(1).norm(dim=1).view(cout, 1, 1), v)
        bound = 1.0 / math.sqrt(cin * k)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))

    def weight(self):
        return weight_norm_apply(self.parametrizations.weight.original1, self.parametrizations.weight.original0)


class _Conv(nn.Module):
    def __init__(self, cin, cout, k, stride=1, pad=0):
        super().__init__()
        self.stride, self.pad = stride, pad
        w = torch.empty(cout, cin, k)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w)
        bound = 1.0 / math.sqrt(cin * k)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


class _Linear(nn.Module):
    """nn.Linear parameters ([out, in]); applied as a K=1 convolution on [B, C, T] activations."""

    def __init__(self, cin, cout):
        super().__init__()
        w = torch.empty(cout, cin)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w)
        bound = 1.0 / math.sqrt(cin)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))

    def w3(self):
        return self.weight.unsqueeze(-1).contiguous()


class _Wrap(nn.Module):
    def __init__(self, name, mod):
        super().__init__()
        self.add_module(name, mod)


# ------------------------------------------------------------------------------------------------ blocks
_SIDE = {}


def _side_streams(device, n):
    """Per-device pool of auxiliary streams for the fork/join sections of the encoder (created once, reused, also under graph capture)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    st = _SIDE.get(key)
    if st is None or len(st) < n:
        st = _SIDE[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return st


class ResBlock1(nn.Module):
    """ttts/vqvae/modules.py:224-318: 3 x [lrelu -> conv(k, d_i) -> lrelu -> conv(k, 1) -> + x]."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([_ConvPWN(channels, channels, kernel_size, dil=d, pad=(kernel_size * d - d) // 2) for d in dilation])
        self.convs2 = nn.ModuleList([_ConvPWN(channels, channels, kernel_size, dil=1, pad=(kernel_size - 1) // 2) for _ in dilation])

    def forward(self, x, out=None, out_scale=1.0, accumulate=False):
        n = len(self.convs1)
        for i, (c1, c2) in enumerate(zip(self.convs1, self.convs2)):
            xt = conv1d(x, c1.weight(), c1.bias, dil=c1.dil, pad=c1.pad, pre_lrelu=True)
            last = i == n - 1
            x = conv1d(xt, c2.weight(), c2.bias, dil=1, pad=c2.pad, pre_lrelu=True, resid=x,
                       out=out if last else None, out_scale=out_scale if last else 1.0, accumulate=accumulate if last else False)
        return x


def _halves(t, H):
    """(t[:H], t[H:]) as contiguous tensors, cached ON the tensor object `t` (address + version checked): the WN res / skip halves are
    sliced out of one weight every call, and a fresh slice would defeat the per-tensor caches (split-bf16 weights) downstream."""
    state = (t.data_ptr(), t._version, tuple(t.shape))
    hit = getattr(t, "_ttts_halves", None)
    if hit is not None and hit[0] == state:
        return hit[1], hit[2]
    lo, hi = t[:H].detach().contiguous(), t[H:].detach().contiguous()
    if not torch.cuda.is_current_stream_capturing():
        t._ttts_halves = (state, lo, hi)
    return lo, hi


class WN(nn.Module):
    """ttts/vqvae/modules.py:136-221 (gated dilated convs with global conditioning)."""

    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0):
        super().__init__()
        self.hidden_channels, self.n_layers, self.gin_channels = hidden_channels, n_layers, gin_channels
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        if gin_channels != 0:
            self.cond_layer = _ConvWN(gin_channels, 2 * hidden_channels * n_layers, 1)
        for i in range(n_layers):
            d = dilation_rate ** i
            self.in_layers.append(_ConvWN(hidden_channels, 2 * hidden_channels, kernel_size, dil=d, pad=int((kernel_size * d - d) / 2)))
            self.res_skip_layers.append(_ConvWN(hidden_channels, 2 * hidden_channels if i < n_layers - 1 else hidden_channels, 1))

    def forward(self, x, x_mask, g=None):
        H = self.hidden_channels
        B, _, T = x.shape
        mask2 = x_mask.reshape(B, T).contiguous()
        output = torch.zeros_like(x)
        gc = None
        if g is not None:
            gc = conv1d(g.contiguous(), self.cond_layer.weight(), self.cond_layer.bias).reshape(B, -1)      # [B, 2*H*n_layers]
        for i in range(self.n_layers):
            il, rs = self.in_layers[i], self.res_skip_layers[i]
            cond = gc[:, i * 2 * H:(i + 1) * 2 * H] if gc is not None else None
            acts = conv1d(x, il.weight(), il.bias, dil=il.dil, pad=il.pad, post=3, cond=cond)
            w = rs.weight()
            if i < self.n_layers - 1:
                w_res, w_skip = _halves(w, H)                      # cached on the tensor: the halves (and their split-bf16 forms) are made once
                b_res, b_skip = _halves(rs.bias, H)
                conv1d(acts, w_skip, b_skip, out=output, accumulate=True)
                x = conv1d(acts, w_res, b_res, resid=x, mask=mask2)
            else:
                conv1d(acts, w, rs.bias, out=output, accumulate=True)
        return output * x_mask


class MelStyleEncoder(nn.Module):
    """ttts/vqvae/modules.py:686-764 (eval semantics: dropouts off)."""

    def __init__(self, n_mel_channels=80, style_hidden=128, style_vector_dim=256, style_kernel_size=5, style_head=2, dropout=0.1):
        super().__init__()
        self.in_dim, self.hidden_dim, self.out_dim = n_mel_channels, style_hidden, style_vector_dim
        self.kernel_size, self.n_head = style_kernel_size, style_head
        H = style_hidden
        spectral = nn.Module()
        spectral.add_module("0", _Wrap("fc", _Linear(self.in_dim, H)))
        spectral.add_module("3", _Wrap("fc", _Linear(H, H)))
        self.spectral = spectral
        temporal = nn.Module()
        for i in range(2):
            temporal.add_module(str(i), _Wrap("conv1", _Wrap("conv", _Conv(H, 2 * H, style_kernel_size, pad=(style_kernel_size - 1) // 2))))
        self.temporal = temporal
        slf = nn.Module()
        slf.w_qs, slf.w_ks, slf.w_vs, slf.fc = _Linear(H, H), _Linear(H, H), _Linear(H, H), _Linear(H, H)
        self.slf_attn = slf
        self.fc = _Wrap("fc", _Linear(H, style_vector_dim))

    def forward(self, x, mask=None):
        """x [B, n_mel, T] (already multiplied by the mask by the caller), mask [B,1,T] float -> [B, out_dim, 1]."""
        lib = L.lib(); _protos(lib)
        B, _, T = x.shape
        lens = mask.reshape(B, T).sum(dim=1).to(torch.int64) if mask is not None else None
        keep = mask.reshape(B, T).contiguous() if mask is not None else None
        s0, s3 = getattr(self.spectral, "0").fc, getattr(self.spectral, "3").fc
        h = conv1d(x.contiguous(), s0.w3(), s0.bias, post=2)
        h = conv1d(h, s3.w3(), s3.bias, post=2)
        for i in range(2):
            c = getattr(self.temporal, str(i)).conv1.conv
            h = conv1d(h, c.weight, c.bias, pad=c.pad, post=1, resid=h)
        if keep is not None:
            h = h * keep[:, None, :]                              # masked_fill(mask, 0) before attention
        a = self.slf_attn
        q = conv1d(h, a.w_qs.w3(), a.w_qs.bias)
        k = conv1d(h, a.w_ks.w3(), a.w_ks.bias)
        v = conv1d(h, a.w_vs.w3(), a.w_vs.bias)
        att = torch.empty_like(q)
        L.check(lib.ttts_mha_small(_p(q), _p(k), _p(v), _p(lens), _p(att), B, self.hidden_dim, T, self.n_head, float(self.hidden_dim ** 0.5),
                                   L.stream_ptr().value), "ttts_mha_small")
        h = conv1d(att, a.fc.w3(), a.fc.bias, resid=h)
        h = conv1d(h, self.fc.fc.w3(), self.fc.fc.bias)
        w = torch.empty(B, self.out_dim, dtype=torch.float32, device=x.device)
        L.check(lib.ttts_masked_mean(_p(h), _p(lens), _p(w), B, self.out_dim, T, L.stream_ptr().value), "ttts_masked_mean")
        return w.unsqueeze(-1)


class _SnakeBeta(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.alpha = nn.Parameter(torch.zeros(c))
        self.beta = nn.Parameter(torch.zeros(c))


class _Filt(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("filter", kaiser_sinc_filter12().view(1, 1, 12))


class _LowPass(nn.Module):
    def __init__(self):
        super().__init__()
        self.lowpass = _Filt()


class _Activation1d(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.act = _SnakeBeta(c)
        self.upsample = _Filt()
        self.downsample = _LowPass()

    def forward(self, x):
        lib = L.lib(); _protos(lib)
        B, C, T = x.shape
        y = torch.empty_like(x)
        L.check(lib.ttts_snake_aa(_p(x), _p(self.act.alpha), _p(self.act.beta), _p(self.upsample.filter), _p(y), B, C, T, L.stream_ptr().value),
                "ttts_snake_aa")
        return y


class PosteriorAudioEncoder(nn.Module):
    """ttts/vqvae/vq2.py:667-745."""

    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0):
        super().__init__()
        self.in_channels, self.out_channels, self.hidden_channels = in_channels, out_channels, hidden_channels
        self.pre = _Conv(in_channels, hidden_channels, 1)
        self.down_pre = _Conv(1, 16, 7, pad=3)
        rates, ksz, ch = [10, 8, 2, 2, 2], [16, 16, 8, 2, 2], [16, 32, 64, 96, 128, 192]
        self.num_kernels = 3
        self.downs = nn.ModuleList([_ConvWN(ch[i], ch[i + 1], k, stride=u, pad=(k - 1) // 2) for i, (u, k) in enumerate(zip(rates, ksz))])
        self.resblocks = nn.ModuleList()
        for i in range(5):
            for k in (3, 7, 11):
                self.resblocks.append(ResBlock1(ch[i + 1], k, (1, 3, 5)))
        self.activation_post = _Activation1d(ch[-1])
        self.conv_post = _Conv(ch[-1], hidden_channels, 7, pad=3)
        self.enc = WN(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.proj = _Conv(hidden_channels * 2, out_channels * 2, 1)

    def forward(self, x, x_audio, x_mask, g=None, eps=None):
        """x: spectrogram [B,1025,T]; x_audio: [B,1,L]; x_mask [B,1,T] float.  Returns (z, m, logs); `eps` replaces the reference's
        torch.randn_like(m) (None = 0, i.e. z = m: the deterministic encode used for extraction / tests)."""
        lib = L.lib(); _protos(lib)
        B, _, T = x.shape
        torch.cuda.nvtx.range_push("ttts.vqenc.posterior_encoder")
        mask2 = x_mask.reshape(B, T).contiguous()
        # Two independent branches meet at `cat`: the spectrogram branch (pre -> 16 gated WN layers, T = 36 frames: small launches) and
        # the waveform branch (strided convs + 15 ResBlocks).  They run on two streams, and inside the waveform branch the three
        # ResBlocks of a level (kernel sizes 3/7/11, same input) run on three streams: streams (and the CUDA graph captured from them)
        # instead of one serial chain of ~350 under-filled launches.  The sums keep the serial order ((r0 + r1) + r2): same bits.
        main = torch.cuda.current_stream()
        side = _side_streams(x.device, 3)
        fork = torch.cuda.Event()
        fork.record(main)
        cat = torch.empty(B, 2 * self.hidden_channels, T, dtype=torch.float32, device=x.device)
        with torch.cuda.stream(side[2]):
            side[2].wait_event(fork)
            if callable(g):                                  # TTTS_ENC_OVERLAP=1: the style encoder runs here, beside the waveform branch
                g = g()
                self.last_g = g
                if not torch.cuda.is_current_stream_capturing():
                    g.record_stream(main)
            h = conv1d(x.contiguous(), self.pre.weight, self.pre.bias, mask=mask2)
            hb = self.enc(h, x_mask, g=g)
            if not torch.cuda.is_current_stream_capturing():
                hb.record_stream(main)                       # consumed on the main stream after the join
        a = conv1d(x_audio.contiguous(), self.down_pre.weight, self.down_pre.bias, pad=3)
        nk = self.num_kernels
        for i in range(5):
            dn = self.downs[i]
            a = conv1d(a, dn.weight(), dn.bias, stride=dn.stride, pad=dn.pad)
            outs = [torch.empty_like(a) for _ in range(nk)]          # allocated on the main stream, written by the branch streams
            ev = torch.cuda.Event()
            ev.record(main)
            for j in range(nk):
                st = main if j == 0 else side[j - 1]
                with torch.cuda.stream(st):
                    if j > 0:
                        st.wait_event(ev)
                    self.resblocks[i * nk + j](a, out=outs[j], out_scale=1.0 / nk, accumulate=False)
            for j in range(1, nk):
                main.wait_stream(side[j - 1])
            xs = outs[0]
            for j in range(1, nk):
                xs = xs.add_(outs[j])
            a = xs
        a = self.activation_post(a)
        assert a.shape[-1] == T, "audio / spectrogram frame mismatch (%d vs %d)" % (a.shape[-1], T)
        cat[:, self.hidden_channels:] = conv1d(a, self.conv_post.weight, self.conv_post.bias, pad=3, mask=mask2)
        main.wait_stream(side[2])
        cat[:, :self.hidden_channels] = hb
        stats = conv1d(cat, self.proj.weight, self.proj.bias, mask=mask2)
        m, logs = torch.split(stats, self.out_channels, dim=1)
        z = torch.empty(B, self.out_channels, T, dtype=torch.float32, device=x.device)
        L.check(lib.ttts_posterior_sample(_p(stats), _p(eps.contiguous()) if eps is not None else None, _p(mask2), _p(z), B, self.out_channels, T,
                                          L.stream_ptr().value), "ttts_posterior_sample")
        torch.cuda.nvtx.range_pop()
        return z, m, logs


class VQEncoder(nn.Module):
    """The encode half of SynthesizerTrn (vq2.py:826-836, 843-852): ref_enc + enc_p + proj + quantizer with the reference's names."""

    def __init__(self, spec_channels=1025, inter_channels=192, hidden_channels=192, gin_channels=512, n_fft=2048, hop=640):
        super().__init__()
        self.n_fft, self.hop = n_fft, hop
        self.enc_p = PosteriorAudioEncoder(spec_channels, inter_channels, hidden_channels, 5, 1, 16, gin_channels=gin_channels)
        self.ref_enc = MelStyleEncoder(spec_channels, style_vector_dim=gin_channels)
        self.quantizer = ResidualVectorQuantizer(dimension=inter_channels, n_q=1, bins=1024)
        self.proj = _Conv(inter_channels, inter_channels, 2, stride=2)

    # convolutions of the extraction / eval forward on the tcgen05 tensor cores with split-bf16 operands (conv1d_tcs: ~1.3e-5 of the fp32
    # reference, codes identical on the golden clips, 2.2x faster).  TTTS_CONV_TC=0 (or `encoder.conv_tc = False`) = exact-fp32 kernels.
    conv_tc = os.environ.get("TTTS_CONV_TC", "1") != "0"

    @torch.no_grad()
    def forward(self, wav, lengths=None, eps=None, sample=False):
        global USE_TC
        prev, USE_TC = USE_TC, bool(self.conv_tc)
        try:
            return self._forward(wav, lengths, eps, sample)
        finally:
            USE_TC = prev

    def _forward(self, wav, lengths=None, eps=None, sample=False):
        """wav [B, L] fp32 (L a multiple of hop).  Returns dict(spec, ge, z, m, logs, x, codes [1,B,N], quantized).
        Posterior noise: the reference draws `randn_like(m)` even in eval (vq2.py:744).  Pass `eps` [B,192,T] to fix it (parity tests),
        `sample=True` to draw it here with torch's generator (reference behaviour), or neither for the deterministic z = m encode
        (extraction default: the same clip always maps to the same codes)."""
        L.require_cuda(wav)
        B = wav.shape[0]
        spec = spectrogram_torch(wav, self.n_fft, self.hop, self.n_fft, center=False)
        T = spec.shape[-1]
        if eps is None and sample:
            eps = torch.randn(B, self.enc_p.out_channels, T, dtype=torch.float32, device=wav.device)
        if lengths is None:
            mask = torch.ones(B, 1, T, dtype=torch.float32, device=wav.device)
        else:
            mask = (torch.arange(T, device=wav.device)[None, :] < lengths[:, None]).float().unsqueeze(1)      # commons.sequence_mask
        if os.environ.get("TTTS_ENC_OVERLAP", "0") == "1":
            # opt-in (TTTS_ENC_OVERLAP=1; measured slower in r2b, 8.9 vs 7.9 ms): only the WN branch needs the style vector, so MelStyleEncoder (~15 small launches, two of them
            # 1025-row reductions) moves onto the WN branch's stream and overlaps the waveform branch instead of preceding both.  Same kernels,
            # same inputs: bit-identical results.
            z, m, logs = self.enc_p(spec, wav.unsqueeze(1), mask, g=lambda: self.ref_enc(spec * mask, mask), eps=eps)
            ge = self.enc_p.last_g                           # joined: enc_p waited for the side stream before returning
        else:
            ge = self.ref_enc(spec * mask, mask)
            z, m, logs = self.enc_p(spec, wav.unsqueeze(1), mask, g=ge, eps=eps)
        x = conv1d(z, self.proj.weight, self.proj.bias, stride=2)
        self.quantizer.eval()
        quantized, codes, commit, _ = self.quantizer(x, layers=[0])
        return dict(spec=spec, ge=ge, z=z, m=m, logs=logs, x=x, codes=codes, quantized=quantized)

    @torch.no_grad()
    def extract_latent(self, wav, y=None, y_lengths=None, sample=False):
        """SynthesizerTrn.extract_latent (vq2.py:912-920): codes [B, n_q, N].  The reference reads an undefined `y_lengths` there (it only
        works through a global); here it is an explicit optional argument (None = all frames valid).  `y` (the linear spectrogram) is
        recomputed from `wav` by the fused STFT kernel, so a caller-supplied `y` is only shape-checked."""
        out = self.forward(wav, lengths=y_lengths, sample=sample)
        if y is not None and tuple(y.shape) != tuple(out["spec"].shape):
            raise ValueError("extract_latent: spectrogram shape %s does not match the waveform (%s)" % (tuple(y.shape), tuple(out["spec"].shape)))
        return out["codes"].transpose(0, 1)

    @torch.no_grad()
    def encode_graphed(self, wav):
        """Extraction fast path: the whole wav -> codes forward (≈ 350 small launches) captured once per input shape in a
        CUDA graph and replayed (streams + graphs instead of a tracing compiler).  Returns codes [1, B, N] (a static buffer)."""
        key = tuple(wav.shape) + (bool(self.conv_tc),)
        st = getattr(self, "_graphs", None)
        if st is None:
            st = self._graphs = {}
        if key not in st:
            static_in = wav.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.forward(static_in)             # warm-up: attribute settings, tensor-map cache, weight-norm cache
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.forward(static_in)
            st[key] = (g, static_in, out)
        g, static_in, out = st[key]
        static_in.copy_(wav)
        g.replay()
        return out["codes"]
