"""CPU emulation of the log-mel backward kernel (tests/emu builds ttts_b200/csrc/stft_bwd.cu for the host) against torch.autograd through a
torch restatement of `mel_spectrogram_torch` (ttts/utils/data_utils.py:106-156: reflect pad, hann STFT, sqrt(re^2 + im^2 + 1e-6), sparse mel
basis, log(clamp(., 1e-5))) -- the op the mel-reconstruction loss of the VQ-VAE-GAN step differentiates through."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu_build import compile_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libstft_bwd_emu.so")
    compile_emu("stft_bwd_emu.cpp", so)
    lib = ctypes.CDLL(so)
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
    lib.ttts_stft_mel_bwd.argtypes = [vp, i32, i32, i32, i32, i32, vp, f32, i32, vp, vp, vp, f32, vp, i32, vp, vp]
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def triangular_bands(n_bins, n_mels, rs):
    """a sparse filterbank in the layout ttts_stft_mel takes: band m covers bins lo[m] .. lo[m] + cnt[m], weights w[off[m] ..]"""
    edges = np.sort(rs.choice(np.arange(1, n_bins - 1), size=n_mels + 2, replace=False))
    lo, off, w, dense = [], [0], [], np.zeros((n_mels, n_bins), np.float32)
    for m in range(n_mels):
        a, c, b = edges[m], edges[m + 1], edges[m + 2]
        ks = np.arange(a, b + 1)
        tri = np.where(ks <= c, (ks - a) / max(c - a, 1), (b - ks) / max(b - c, 1)).astype(np.float32) + 0.01
        lo.append(a); w.extend(tri.tolist()); off.append(off[-1] + len(ks))
        dense[m, a:b + 1] = tri
    return np.array(lo, np.int32), np.array(off, np.int32), np.array(w, np.float32), dense


@pytest.mark.parametrize("n_fft,hop,L,n_mels", [(64, 20, 200, 8), (256, 80, 640, 16)])
def test_log_mel_backward(emu, n_fft, hop, L, n_mels):
    rs = np.random.RandomState(n_fft)
    B, pad = 2, (n_fft - hop) // 2
    lo, off, w, dense = triangular_bands(n_fft // 2 + 1, n_mels, rs)
    wav = torch.tensor(np.clip(0.3 * rs.standard_normal((B, L)), -1, 1).astype(np.float32), requires_grad=True)
    window = torch.hann_window(n_fft)
    xp = F.pad(wav.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(xp, n_fft, hop_length=hop, win_length=n_fft, window=window, center=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-6)
    mel = torch.log(torch.clamp(torch.tensor(dense) @ mag, min=1e-5))
    Fr = mel.shape[-1]
    dlog = torch.tensor(rs.standard_normal((B, n_mels, Fr)).astype(np.float32))
    mel.backward(dlog)
    dwav = torch.zeros(B, L)
    lo_t, off_t, w_t = torch.tensor(lo), torch.tensor(off), torch.tensor(w)
    rc = emu.ttts_stft_mel_bwd(wav.detach().data_ptr(), B, L, n_fft, hop, pad, window.data_ptr(), 1e-6, n_mels, lo_t.data_ptr(), off_t.data_ptr(),
                               w_t.data_ptr(), 1e-5, dlog.data_ptr(), Fr, dwav.data_ptr(), None)
    assert rc == 0, emu.emu_last_error()
    assert float((dwav - wav.grad).norm()) <= 2e-4 * float(wav.grad.norm()), float((dwav - wav.grad).norm() / wav.grad.norm())
    # frame count mismatches are reported
    assert emu.ttts_stft_mel_bwd(wav.detach().data_ptr(), B, L, n_fft, hop, pad, window.data_ptr(), 1e-6, n_mels, lo_t.data_ptr(), off_t.data_ptr(),
                                 w_t.data_ptr(), 1e-5, dlog.data_ptr(), Fr + 1, dwav.data_ptr(), None) != 0
