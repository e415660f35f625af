"""First timing of the generator half of the VQ-VAE-GAN train step (BASELINE.json config 4: enc + VQ + dec + disc) on the
training tape (ttts_b200/vqvae/train_step.py) with the reference's segment size (20 480 samples = 32 frames) at batch B clips of 23 040
samples.  Correctness-first kernels (no pipelining, grouped convolutions per group): the number is a starting point, not a claim.
Prints one JSON object.   python tools/vqvae_step_bench.py [B] [iters]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import torch
    import make_golden as MG
    from ttts_b200 import _lib as L
    from ttts_b200.vqvae.mel import spectrogram_torch
    from ttts_b200.vqvae.train_encoder import CudaKernels
    from ttts_b200.vqvae.train_step import GeneratorStep

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda")
    G, D = MG.step_params()
    G = {k: v.to(dev) for k, v in G.items()}
    D = {k: v.to(dev) for k, v in D.items()}
    g = torch.Generator(device="cuda").manual_seed(1234)
    wav = torch.clamp(0.1 * torch.randn(B, 23040, device=dev, generator=g), -1, 1)
    lengths = torch.full((B,), 36, dtype=torch.int64, device=dev)
    text = torch.randint(0, 256, (B, 40), device=dev, generator=g)
    text_lengths = torch.full((B,), 40, dtype=torch.int64, device=dev)
    E = torch.randn(1024, 192, device=dev, generator=g)
    eps_p, eps_q = torch.randn(B, 192, 36, device=dev, generator=g), torch.randn(B, 192, 36, device=dev, generator=g)
    ids = [int(i) % 5 for i in range(B)]
    lib = L.lib()
    lib.ttts_launch_count.restype = __import__("ctypes").c_ulonglong
    K = CudaKernels()
    times, launches = [], 0
    for it in range(iters + 1):
        spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
        torch.cuda.synchronize()
        l0, t0 = lib.ttts_launch_count(), time.perf_counter()
        step = GeneratorStep(K, G, D)
        out = step.forward(wav, spec, lengths, text, text_lengths, E, eps_p, eps_q, ids, 32)
        grads = step.backward()
        torch.cuda.synchronize()
        if it > 0:
            times.append(time.perf_counter() - t0)
            launches = lib.ttts_launch_count() - l0
    ms = 1e3 * sorted(times)[len(times) // 2]
    print(json.dumps({"workload": "VQ-VAE-GAN generator step (fwd + bwd, no optimizer), segment 20480", "B": B, "ms_per_step": ms,
                      "samples_per_s": B * 23040 / ms * 1e3, "kernel_launches": int(launches), "n_grad_tensors": len(grads),
                      "losses": {k: float(out[k].v) for k in ("loss_gen", "loss_fm", "loss_mel", "kl_ssl", "loss_kl")},
                      "timing": "host wall clock around a synchronised step (hundreds of small launches: launch-bound by construction)"}))


if __name__ == "__main__":
    main()
