"""STFT / mel front end -- drop-ins for `spectrogram_torch`, `spec_to_mel_torch`, `mel_spectrogram_torch`
(ttts/utils/data_utils.py:52-156) and `MelSpectrogramFeatures` (ttts/vocoder/feature_extractors.py:28-49), backed by the
batched shared-memory rFFT + sparse-mel kernel in csrc/stft.cu.  Same signatures, same output layouts ([B, bins, F] /
[B, n_mels, F], fp32).  No CPU fallback.

Host-side constants (hann window, FFT twiddles, mel filterbanks) are computed once per (config, device) in float64 and
cached, like the reference's module-level `hann_window` / `mel_basis` dicts (data_utils.py:48-49).  The Slaney filterbank
restates librosa.filters.mel (the reference's unpinned third-party dependency); the HTK one restates
torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk").
"""
import ctypes
import math

import numpy as np
import torch
from torch import nn

from .. import _lib as L

_cache = {}


def _protos(lib):
    if getattr(lib, "_stft_protos", False):
        return
    lib.ttts_stft_mel.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    lib.ttts_logmel.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    lib._stft_protos = True


# ------------------------------------------------------------------------------------------------ host constants
def _hz_to_mel_slaney(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney') -> [n_mels, n_fft/2+1] float32."""
    fmax = fmax or sr / 2.0
    fft_freqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    pts = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(pts)
    ramps = pts[:, None] - fft_freqs[None, :]
    w = np.maximum(0, np.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]))
    w *= (2.0 / (pts[2:n_mels + 2] - pts[:n_mels]))[:, None]
    return w.astype(np.float32)


def htk_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk').T -> [n_mels, n_fft/2+1] float32."""
    fmax = fmax or sr / 2.0
    all_freqs = np.linspace(0, sr // 2, n_fft // 2 + 1)
    m_pts = np.linspace(2595.0 * np.log10(1.0 + fmin / 700.0), 2595.0 * np.log10(1.0 + fmax / 700.0), n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    fb = np.maximum(0, np.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    return fb.T.astype(np.float32)


def sparsify(basis):
    """Each mel band is one contiguous run of non-zero bins: (band_lo[n_mels], band_off[n_mels+1], weights)."""
    lo, off, w = [], [0], []
    for row in basis:
        nz = np.nonzero(row)[0]
        if len(nz) == 0:
            lo.append(0)
            off.append(off[-1])
            continue
        a, b = int(nz[0]), int(nz[-1]) + 1
        lo.append(a)
        w.append(row[a:b])
        off.append(off[-1] + (b - a))
    wcat = np.concatenate(w).astype(np.float32) if w else np.zeros(1, np.float32)
    return np.asarray(lo, np.int32), np.asarray(off, np.int32), wcat


def _stft_consts(n_fft, win_size, device):
    key = ("stft", n_fft, win_size, str(device))
    if key not in _cache:
        if win_size != n_fft:
            raise NotImplementedError("win_size != n_fft is not used by the reference's front end")
        n = np.arange(win_size)
        window = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_size)).astype(np.float32)       # torch.hann_window (periodic)
        k = np.arange(n_fft // 2 + 1)
        tw = np.stack([np.cos(2.0 * np.pi * k / n_fft), -np.sin(2.0 * np.pi * k / n_fft)], axis=1).astype(np.float32)
        _cache[key] = (torch.from_numpy(window).to(device), torch.from_numpy(tw).contiguous().to(device))
    return _cache[key]


def _mel_consts(kind, sr, n_fft, n_mels, fmin, fmax, device):
    key = ("mel", kind, sr, n_fft, n_mels, fmin, fmax, str(device))
    if key not in _cache:
        basis = slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax) if kind == "slaney" else htk_mel_basis(sr, n_fft, n_mels, fmin, fmax)
        lo, off, w = sparsify(basis)
        _cache[key] = (torch.from_numpy(lo).to(device), torch.from_numpy(off).to(device), torch.from_numpy(w).to(device))
    return _cache[key]


def _run(y, n_fft, hop, pad, eps_inside, want_spec, mel, log_floor):
    lib = L.lib(); _protos(lib)
    L.require_cuda(y)
    squeeze = y.dim() == 1
    y2 = (y[None] if squeeze else y).contiguous().float()
    B, Lw = y2.shape
    F = 1 + (Lw + 2 * pad - n_fft) // hop
    window, tw = _stft_consts(n_fft, n_fft, y2.device)
    spec = torch.empty(B, n_fft // 2 + 1, F, dtype=torch.float32, device=y2.device) if want_spec else None
    melo = None
    lo = off = w = None
    n_mels = 0
    if mel is not None:
        lo, off, w = mel
        n_mels = lo.numel()
        melo = torch.empty(B, n_mels, F, dtype=torch.float32, device=y2.device)
    L.check(lib.ttts_stft_mel(y2.data_ptr(), B, Lw, n_fft, hop, pad, window.data_ptr(), tw.data_ptr(), float(eps_inside),
                              spec.data_ptr() if spec is not None else None, n_mels,
                              lo.data_ptr() if lo is not None else None, off.data_ptr() if off is not None else None,
                              w.data_ptr() if w is not None else None, float(log_floor),
                              melo.data_ptr() if melo is not None else None, F, L.stream_ptr().value), "ttts_stft_mel")
    if squeeze:
        spec = spec[0] if spec is not None else None
        melo = melo[0] if melo is not None else None
    return spec, melo


# ------------------------------------------------------------------------------------------------ reference API
def spectrogram_torch(y, n_fft, hop_size, win_size, center=False):
    """ttts/utils/data_utils.py:52-87: reflect-pad (n_fft-hop)/2, hann(win), stft(center=False), sqrt(re^2+im^2+1e-6)."""
    if center:
        raise NotImplementedError("the reference only calls spectrogram_torch with center=False")
    if win_size != n_fft:
        raise NotImplementedError("win_size != n_fft")
    spec, _ = _run(y, n_fft, hop_size, int((n_fft - hop_size) / 2), 1e-6, True, None, 0.0)
    return spec


def spec_to_mel_torch(spec, n_fft, num_mels, sampling_rate, fmin, fmax):
    """ttts/utils/data_utils.py:90-103: log(clamp(mel_basis @ spec, 1e-5))."""
    lib = L.lib(); _protos(lib)
    L.require_cuda(spec)
    s = spec.contiguous().float()
    squeeze = s.dim() == 2
    if squeeze:
        s = s[None]
    B, bins, F = s.shape
    lo, off, w = _mel_consts("slaney", sampling_rate, n_fft, num_mels, fmin, fmax, s.device)
    out = torch.empty(B, num_mels, F, dtype=torch.float32, device=s.device)
    L.check(lib.ttts_logmel(s.data_ptr(), B, bins, F, num_mels, lo.data_ptr(), off.data_ptr(), w.data_ptr(), 1e-5, out.data_ptr(),
                            L.stream_ptr().value), "ttts_logmel")
    return out[0] if squeeze else out


def mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """ttts/utils/data_utils.py:106-156 (spectrogram_torch + spec_to_mel_torch fused in one kernel)."""
    if center or win_size != n_fft:
        raise NotImplementedError("center=True / win_size != n_fft are not used by the reference")
    mel = _mel_consts("slaney", sampling_rate, n_fft, num_mels, fmin, fmax, y.device)
    _, m = _run(y, n_fft, hop_size, int((n_fft - hop_size) / 2), 1e-6, False, mel, 1e-5)
    return m


class MelSpectrogramFeatures(nn.Module):
    """ttts/vocoder/feature_extractors.py:28-49: torchaudio MelSpectrogram(sr, n_fft, hop, n_mels, center=True, power=1)
    (HTK mels, no norm) followed by safe_log = log(clip(x, 1e-7))."""

    def __init__(self, sample_rate=24000, n_fft=1024, hop_length=256, n_mels=100, padding="center"):
        super().__init__()
        if padding not in ["center", "same"]:
            raise ValueError("Padding must be 'center' or 'same'.")
        if padding == "same":
            raise NotImplementedError("padding='same' is not used by the reference's pipelines")
        self.padding = padding
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels

    def forward(self, audio, **kwargs):
        mel = _mel_consts("htk", self.sample_rate, self.n_fft, self.n_mels, 0.0, None, audio.device)
        _, m = _run(audio, self.n_fft, self.hop_length, self.n_fft // 2, 0.0, False, mel, 1e-7)
        return m
