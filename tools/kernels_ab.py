"""A/B harness for the VQ-encode front-end kernels: times STFT+mel (v2 front end, 4096 clips; 24 kHz front end, 4096 clips) and the VQ
argmin (N = 1 152 and 2^20), prints digests of the integer outputs and, when a previous run left /tmp/kernels_ab.npz, the largest
deviation from it.  Kernel variants are chosen by environment (TTTS_STFT_V1=1, TTTS_VQ_V1=1): run once per variant.
ONLY=stft|vq|mel24 restricts the run (for ncu captures); ITERS overrides the repeat count."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from ttts_b200.vqvae.mel import mel_spectrogram_torch, spectrogram_torch, MelSpectrogramFeatures
from ttts_b200.vqvae.quantize import vq_lookup

only = os.environ.get("ONLY", "")
iters = int(os.environ.get("ITERS", "5"))
tag = "stft_v1=%s vq_v1=%s" % (os.environ.get("TTTS_STFT_V1", "0"), os.environ.get("TTTS_VQ_V1", "0"))
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(1234)
HBM = 6538.0
FP32 = 72.0     # TFLOP/s, FMA = 2 FLOP (148 SMs x 128 lanes x 2 x ~1.9 GHz)


def t(fn, n=iters, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


keep = {}
if only in ("", "stft"):
    Lw, F = 23040, 36
    big = torch.clamp(0.1 * torch.randn(4096, Lw, device=dev, generator=g), -1, 1)
    ms = t(lambda: mel_spectrogram_torch(big, 2048, 128, 32000, 640, 2048, 0, None))
    mss = t(lambda: spectrogram_torch(big, 2048, 640, 2048))
    nfr = 4096 * F
    flop = nfr * (5 * 1024 * 10 + 8 * 1025 + 2 * 2200)          # 1024-point complex FFT + unpack/magnitude + sparse mel
    print("%s | stft+mel 4096 clips: %.3f ms  %.1f Mframes/s  %.0f GB/s (%.3f of HBM)  %.2f TFLOP/s (%.3f of fp32 FMA peak) | spectrogram only: %.3f ms %.0f GB/s (%.3f of HBM)" % (
        tag, ms, nfr / ms / 1e3, (big.numel() * 4 + nfr * 128 * 4) / ms / 1e6, (big.numel() * 4 + nfr * 128 * 4) / ms / 1e6 / HBM, flop / ms / 1e9, flop / ms / 1e9 / FP32,
        mss, (big.numel() * 4 + nfr * 1025 * 4) / mss / 1e6, (big.numel() * 4 + nfr * 1025 * 4) / mss / 1e6 / HBM), flush=True)
    keep["mel"] = mel_spectrogram_torch(big[:8], 2048, 128, 32000, 640, 2048, 0, None).cpu().numpy()
    keep["spec"] = spectrogram_torch(big[:8], 2048, 640, 2048).cpu().numpy()
    del big
if only in ("", "mel24"):
    m24 = MelSpectrogramFeatures()
    w24 = torch.clamp(0.1 * torch.randn(4096, 24000, device=dev, generator=g), -1, 1)
    ms24 = t(lambda: m24(w24))
    print("%s | mel24k 4096 clips: %.3f ms  %.1f Mframes/s  %.0f Msamples/s" % (tag, ms24, 4096 * 94 / ms24 / 1e3, w24.numel() / ms24 / 1e3), flush=True)
    keep["mel24"] = m24(w24[:8]).cpu().numpy()
    del w24
if only in ("", "vq"):
    E = torch.randn(1024, 192, device=dev, generator=g)
    for N in (1152, 1 << 20):
        x = torch.randn(N, 192, device=dev, generator=g)
        msv = t(lambda: vq_lookup(x, E, False), n=iters if N > 10000 else 20)
        out = vq_lookup(x, E, False)
        codes = out[0] if isinstance(out, (tuple, list)) else out
        dig = hashlib.sha1(codes.cpu().numpy().tobytes()).hexdigest()[:12]
        print("%s | vq argmin N=%d: %.3f ms  %.2f TFLOP/s fp32 (%.3f of FMA peak)  %.1f GB/s  codes digest %s" % (
            tag, N, msv, N * 393216 / msv / 1e9, N * 393216 / msv / 1e9 / FP32, N * 1544 / msv / 1e6, dig), flush=True)
        del x
prev = "/tmp/kernels_ab.npz"
if os.path.exists(prev):
    old = np.load(prev)
    for k, v in keep.items():
        if k in old.files:
            print("%s | max |%s - previous run| = %.3e (max |value| %.3f)" % (tag, k, float(np.abs(v - old[k]).max()), float(np.abs(v).max())), flush=True)
else:
    np.savez(prev, **keep)
if only == "enc":
    # whole encode (64 clips): time + digests of z and codes.  TTTS_CONV_DIRECT=0 / TTTS_CONV_PIPE=0 select the older conv kernels,
    # which accumulate in the same order: the digests must not change.
    from ttts_b200.vqvae.encoder import VQEncoder
    torch.manual_seed(0)
    wav = torch.clamp(0.1 * torch.randn(64, 23040, device=dev, generator=g), -1, 1)
    enc = VQEncoder().to(dev).eval()
    cb = enc.quantizer.vq.layers[0]._codebook
    cb.embed.copy_(torch.randn(1024, 192, device=dev, generator=g)); cb.inited.fill_(1)
    ms = t(lambda: enc(wav), n=10)
    out = enc(wav)
    dz = hashlib.sha1(out["z"].cpu().numpy().tobytes()).hexdigest()[:12]
    dc = hashlib.sha1(out["codes"].cpu().numpy().tobytes()).hexdigest()[:12]
    print("conv_direct=%s conv_pipe=%s | encode 64 clips: %.3f ms  %.1f Msamples/s  z digest %s codes digest %s" % (
        os.environ.get("TTTS_CONV_DIRECT", "1"), os.environ.get("TTTS_CONV_PIPE", "1"), ms, wav.numel() / ms / 1e3, dz, dc), flush=True)
