"""VQ-encode Msamples/s (BASELINE.json metric 2): wav -> STFT -> ref_enc -> enc_p -> proj -> VQ codes at B = 64 clips of
23 040 samples, plus the bandwidth-bound kernels alone (STFT+mel GB/s, VQ argmin at N = 1 152 and N = 2^20).  Prints one JSON object.
Imported by bench.py (key "vq_encode"); `python tools/vq_encode_bench.py [--cpu]` runs it stand-alone (--cpu adds the oracle timing)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(hbm_gbs=6570.3, with_cpu=False, bf16_tflops=1638.6):
    import numpy as np
    import torch
    from ttts_b200.vqvae.encoder import VQEncoder
    from ttts_b200.vqvae.mel import mel_spectrogram_torch, spectrogram_torch, MelSpectrogramFeatures
    from ttts_b200.vqvae.quantize import vq_lookup

    dev = torch.device("cuda")
    torch.manual_seed(0)
    B, Lw = 64, 23040
    g = torch.Generator(device="cuda").manual_seed(1234)
    wav = torch.clamp(0.1 * torch.randn(B, Lw, device=dev, generator=g), -1, 1)
    enc = VQEncoder().to(dev).eval()
    cb = enc.quantizer.vq.layers[0]._codebook
    cb.embed.copy_(torch.randn(1024, 192, device=dev, generator=g)); cb.inited.fill_(1)
    out = {}
    ms = _time(lambda: enc(wav), iters=10)
    out["encode_ms_per_batch"] = ms
    out["encode_msamples_per_s"] = B * Lw / ms / 1e3
    try:
        msg = _time(lambda: enc.encode_graphed(wav), iters=20)
        out["encode_graphed_ms_per_batch"] = msg
        out["encode_graphed_msamples_per_s"] = B * Lw / msg / 1e3
        assert torch.equal(enc.encode_graphed(wav), enc(wav)["codes"])
    except Exception as e:
        out["encode_graphed_error"] = repr(e)[:200]
    out["batch"] = B
    out["samples_per_clip"] = Lw
    # the convolution stack (90 % of the encode): which kernels run it, and the roofline of the tensor-core convolution
    from ttts_b200.vqvae import encoder as ENC
    from ttts_b200 import _lib as L
    import ctypes
    out["conv_path"] = "conv1d_tcs (split-bf16 tcgen05; default)" if enc.conv_tc else "fp32 CUDA-core kernels (TTTS_CONV_TC=0)"
    if enc.conv_tc:
        lib = L.lib()
        enc(wav); torch.cuda.synchronize()
        lib.ttts_prof_gemm_enable(5)
        enc(wav)
        ms_k, fl_k, n_k = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.ttts_prof_gemm_read(ctypes.byref(ms_k), ctypes.byref(fl_k), ctypes.byref(n_k))
        lib.ttts_prof_gemm_enable(0)
        if ms_k.value > 0:
            tf = fl_k.value / (ms_k.value * 1e-3) / 1e12
            out["conv"] = {"kernel": "conv1d_tcs_kernel", "launches": int(n_k.value), "ms_sum": ms_k.value, "gflop": fl_k.value / 1e9,
                           "tflops": tf, "frac_tensor": tf / (bf16_tflops / 3.0), "frac_fp32": tf / (148 * 128 * 2 * 1.965e-3),
                           "tensor_roof": "dense bf16 peak / 3 (three bf16 products per fp32 product)", "note": "sum of the kernels' own durations (CUDA events per launch, streams serialised by the events); share of the fp32-path encode they replace: ~75 % of its FLOPs"}
    # end to end with HOST buffers, as an extraction script calls it: pinned waveforms -> device, encode, codes -> host, every batch
    try:
        hw = wav.cpu().pin_memory()

        def e2e():
            return enc(hw.to(dev, non_blocking=True))["codes"].cpu()
        assert torch.equal(e2e(), enc(wav)["codes"].cpu())
        mse = _time(e2e, iters=10)
        out["encode_e2e"] = {"ms_per_batch": mse, "msamples_per_s": B * Lw / mse / 1e3, "h2d_bytes_per_batch": hw.numel() * 4,
                             "d2h_bytes_per_batch": int(enc(wav)["codes"].numel()) * 8}
    except Exception as e:
        out["encode_e2e_error"] = repr(e)[:200]
    # STFT + mel (v2 front end): per frame 640*4 B in, (1025 + 128)*4 B out
    F = 36
    ms_spec = _time(lambda: spectrogram_torch(wav, 2048, 640, 2048), iters=50)
    ms_mel = _time(lambda: mel_spectrogram_torch(wav, 2048, 128, 32000, 640, 2048, 0, None), iters=50)
    out["stft_spec"] = {"ms": ms_spec, "gbs": B * F * (640 * 4 + 1025 * 4) / ms_spec / 1e6, "frac_hbm": B * F * (640 * 4 + 1025 * 4) / ms_spec / 1e6 / hbm_gbs}
    out["stft_mel"] = {"ms": ms_mel, "gbs": B * F * (640 * 4 + 128 * 4) / ms_mel / 1e6}
    big = torch.clamp(0.1 * torch.randn(4096, Lw, device=dev, generator=g), -1, 1)        # 94 M samples: far beyond L2
    ms_big = _time(lambda: mel_spectrogram_torch(big, 2048, 128, 32000, 640, 2048, 0, None), iters=5, warm=1)
    bytes_big = big.numel() * 4 + 4096 * F * 128 * 4
    out["stft_mel_4096clips"] = {"ms": ms_big, "gbs": bytes_big / ms_big / 1e6, "frac_hbm": bytes_big / ms_big / 1e6 / hbm_gbs,
                                 "msamples_per_s": big.numel() / ms_big / 1e3}
    del big
    m24 = MelSpectrogramFeatures()
    w24 = torch.clamp(0.1 * torch.randn(64, 24000, device=dev, generator=g), -1, 1)
    ms24 = _time(lambda: m24(w24), iters=50)
    out["mel24k_64clips"] = {"ms": ms24, "msamples_per_s": w24.numel() / ms24 / 1e3}
    # VQ argmin: 393 216 FLOP and 1 544 B per vector
    E = cb.embed
    for N in (1152, 1 << 20):
        x = torch.randn(N, 192, device=dev, generator=g)
        msv = _time(lambda: vq_lookup(x, E, False), iters=20 if N < 10000 else 5, warm=2)
        tc_path = N >= 4096 and os.environ.get("TTTS_VQ_TC", "1") != "0"
        ent = {"ms": msv, "gbs": N * 1544 / msv / 1e6, "frac_hbm": N * 1544 / msv / 1e6 / hbm_gbs,
               "path": "tcgen05 split-bf16 scores + exact fp32 re-check of the candidates (vq_tc_*)" if tc_path else "exact fp32 FMA sweep (vq_argmin_pipe_kernel)"}
        if tc_path:
            ent["tflops_equiv"] = N * 393216 / msv / 1e9           # 2 K D FLOP per vector, as the fp32 kernel would spend them
            ent["frac_tensor"] = 3 * N * 393216 / msv / 1e9 / bf16_tflops      # three bf16 products per fp32 product
        else:
            ent["tflops_fp32"] = N * 393216 / msv / 1e9
            ent["frac_fp32_fma_peak_72tf"] = N * 393216 / msv / 1e9 / 72.0
        out["vq_argmin_N%d" % N] = ent
    if with_cpu:
        from oracle import encoder_oracle as EO
        from oracle import vq_mel_oracle as V
        nthr = min(32, os.cpu_count() or 1)       # small convs: the full thread pool of a many-core host is slower (see bench.pick_cpu_threads)
        torch.set_num_threads(nthr)
        P = EO.init_params(seed=5)
        wc = wav[:8].cpu()
        spec = torch.tensor(V.spectrogram(wc.numpy()))
        Ec = E.cpu().numpy()
        with torch.no_grad():
            EO.encode(P, spec, wc, codebook=Ec)
            t0 = time.perf_counter()
            spec = torch.tensor(V.spectrogram(wc.numpy()))
            EO.encode(P, spec, wc, codebook=Ec)
            dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"msamples_per_s": 8 * Lw / dt / 1e6, "clips": 8, "cores": nthr, "kind": "port"}
    return out


if __name__ == "__main__":
    print(json.dumps(run(with_cpu="--cpu" in sys.argv)))
