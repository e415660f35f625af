"""CPU ORACLE (test infrastructure, never the product path) for the VQ-VAE encode front end: the RVQ (n_q=1) codebook lookup
with its training-time EMA update, and the STFT / mel feature extractors.  numpy only.

Parity status: the reference has no tests ("parity unpinned"); this file is pinned against golden vectors produced by
running the REAL reference modules (tests/golden/make_golden.py -> tests/golden/vq.npz, mel.npz), see tests/test_oracle_golden_vq_mel.py.

Reference citations (under /root/reference):
  quantize ................ ttts/vqvae/core_vq.py:174-182   dist = -(|x|^2 - 2 x E^T + |e|^2), argmax (first max wins)
  forward / EMA ........... ttts/vqvae/core_vq.py:205-230, 46-51 (ema_inplace, laplace_smoothing)
  straight-through, commit  ttts/vqvae/core_vq.py:303-322
  RVQ wrapper ............. ttts/vqvae/core_vq.py:336-374 ; ttts/vqvae/quantize.py:70-118
  spectrogram_torch ....... ttts/utils/data_utils.py:52-87
  spec_to_mel_torch ....... ttts/utils/data_utils.py:90-103 (+ spectral_normalize_torch: log(clamp(x, 1e-5)))
  mel_spectrogram_torch ... ttts/utils/data_utils.py:106-156
  MelSpectrogramFeatures .. ttts/vocoder/feature_extractors.py:28-49 (torchaudio MelSpectrogram 24k/1024/256/100, HTK, power=1,
                            center=True reflect) ; safe_log ttts/vocoder/modules.py:194-205 : log(clip(x, 1e-7))
  mel bases ............... librosa.filters.mel (Slaney scale + Slaney norm; third-party, unpinned, absent here) and
                            torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk')
"""
import numpy as np


# ------------------------------------------------------------------------------------------------ VQ
def vq_quantize(x, E):
    """x [N,D] fp32, E [K,D] fp32 -> indices int64 [N].  Same association order as the reference expression."""
    x = x.astype(np.float32); E = E.astype(np.float32)
    xx = (x ** 2).sum(1, keepdims=True, dtype=np.float32)
    ee = (E ** 2).sum(1, dtype=np.float32)[None, :]
    dist = -((xx - np.float32(2.0) * (x @ E.T)) + ee)
    return dist.argmax(axis=1).astype(np.int64)       # numpy argmax = first maximal index, like torch.max


def vq_margin(x, E, idx):
    """fp64 top-2 margin of each row, relative to |x||e|: rows with a tiny margin may legitimately flip between
    implementations that sum the 192-term dot product in a different order (SURVEY.md section 7 'hard parts')."""
    x = x.astype(np.float64); E = E.astype(np.float64)
    d = ((x ** 2).sum(1)[:, None] - 2 * x @ E.T + (E ** 2).sum(1)[None, :])
    part = np.partition(d, 1, axis=1)
    gap = part[:, 1] - part[:, 0]
    scale = np.linalg.norm(x, axis=1) * np.linalg.norm(E[idx], axis=1) + 1e-30
    return gap / scale


def rvq_forward(x_bdn, E, cluster_size, embed_avg, training, decay=0.99, eps=1e-5):
    """ResidualVectorQuantizer(n_q=1).forward on x [B,D,N] with an initialised codebook (inited=1).
    Returns dict(quantized [B,D,N], codes [1,B,N], commit (scalar), and the updated buffers when training).
    Dead-code expiry / k-means init consume RNG in the reference and are handled by the host module, not here."""
    B, D, N = x_bdn.shape
    x = np.ascontiguousarray(x_bdn.transpose(0, 2, 1)).reshape(B * N, D).astype(np.float32)
    idx = vq_quantize(x, E)
    q = E[idx].astype(np.float32)
    out = {"codes": idx.reshape(1, B, N)}
    if training:
        K = E.shape[0]
        hist = np.bincount(idx, minlength=K).astype(np.float32)
        cs = (cluster_size * np.float32(decay) + np.float32(1 - decay) * hist).astype(np.float32)
        onehot = np.zeros((B * N, K), np.float32); onehot[np.arange(B * N), idx] = 1
        embed_sum = (x.T @ onehot).T
        ea = (embed_avg * np.float32(decay) + np.float32(1 - decay) * embed_sum).astype(np.float32)
        smoothed = (cs + np.float32(eps)) / (cs.sum(dtype=np.float32) + np.float32(K * eps)) * cs.sum(dtype=np.float32)
        out.update(cluster_size=cs, embed_avg=ea, embed=(ea / smoothed[:, None]).astype(np.float32))
        st = x + (q - x)                      # straight-through value (Appendix E #10)
        out["commit"] = np.float32(((q - x) ** 2).mean(dtype=np.float32))
        out["quantized"] = st.reshape(B, N, D).transpose(0, 2, 1)
    else:
        out["commit"] = np.float32(0.0)
        out["quantized"] = q.reshape(B, N, D).transpose(0, 2, 1)
    return out


def rvq_backward(x_bdn, q_bdn_value, dquantized, dcommit):
    """Gradient wrt x of (quantized, commit): straight-through passes dquantized; commit = mean((q-x)^2) -> 2(x-q)/(N*D)."""
    B, D, N = x_bdn.shape
    # q (not the straight-through value) is needed: recover from the caller
    return dquantized + dcommit * 2.0 * (x_bdn - q_bdn_value) / np.float32(B * N * D)


# ------------------------------------------------------------------------------------------------ STFT / mel
def hann_periodic(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))


def _reflect_pad(y, p):
    return np.concatenate([y[:, 1:p + 1][:, ::-1], y, y[:, -p - 1:-1][:, ::-1]], axis=1)


def stft_mag(y, n_fft, hop, win, center, eps_inside, power1=False):
    """|STFT| frames [B, n_fft/2+1, F].  center=False: reference pre-pads (n_fft-hop)/2 reflect (data_utils.py:66-71);
    center=True: torch.stft-style n_fft/2 reflect pad."""
    y = np.asarray(y, np.float64)
    p = (n_fft - hop) // 2 if not center else n_fft // 2
    y = _reflect_pad(y, p)
    n_frames = 1 + (y.shape[1] - n_fft) // hop
    w = hann_periodic(win)
    frames = np.stack([y[:, i * hop:i * hop + n_fft] for i in range(n_frames)], axis=1) * w[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)                       # [B, F, bins]
    mag2 = spec.real ** 2 + spec.imag ** 2
    mag = np.sqrt(mag2 + eps_inside)
    return mag.transpose(0, 2, 1)


def spectrogram(y, n_fft=2048, hop=640, win=2048):
    """spectrogram_torch(y, 2048, 640, 2048, center=False) -> sqrt(re^2 + im^2 + 1e-6)."""
    return stft_mag(y, n_fft, hop, win, center=False, eps_inside=1e-6).astype(np.float32)


def _hz_to_mel_slaney(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis_slaney(sr=32000, n_fft=2048, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(htk=False, norm='slaney') restated: triangular filters on the Slaney mel scale, area-normalised."""
    fmax = fmax or sr / 2.0
    fft_freqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    mel_pts = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


def mel_basis_htk(sr=24000, n_fft=1024, n_mels=100, fmin=0.0, fmax=None):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') transposed to [n_mels, bins]."""
    fmax = fmax or sr / 2.0
    all_freqs = np.linspace(0, sr // 2, n_fft // 2 + 1)
    m_min = 2595.0 * np.log10(1.0 + fmin / 700.0)
    m_max = 2595.0 * np.log10(1.0 + fmax / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0, np.minimum(down, up))
    return fb.T.astype(np.float32)


def spec_to_mel(spec, basis=None):
    """spec_to_mel_torch(spec, 2048, 128, 32000, 0, None): log(clamp(basis @ spec, 1e-5))."""
    basis = mel_basis_slaney() if basis is None else basis
    m = np.einsum("mk,bkf->bmf", basis.astype(np.float64), spec.astype(np.float64))
    return np.log(np.maximum(m, 1e-5)).astype(np.float32)


def mel_spectrogram(y, basis=None):
    return spec_to_mel(spectrogram(y), basis)


def mel_features_24k(y, basis=None):
    """MelSpectrogramFeatures()(y): |STFT|(1024, hop 256, center reflect), HTK mel 100, log(clip(., 1e-7))."""
    basis = mel_basis_htk() if basis is None else basis
    mag = stft_mag(y, 1024, 256, 1024, center=True, eps_inside=0.0)
    m = np.einsum("mk,bkf->bmf", basis.astype(np.float64), mag)
    return np.log(np.maximum(m, 1e-7)).astype(np.float32)
