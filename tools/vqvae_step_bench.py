"""BASELINE.json config 4: one VQ-VAE-GAN train step (enc + VQ + dec + disc; ttts/vqvae/train.py:330-406) at batch B clips of 23 040 samples,
the reference's segment size (20 480 samples = 32 frames), through `ttts_b200.vqvae.train_step.TrainStep.step` -- synthesis -> discriminator
loss -> optim_d -> adversarial + feature losses through the updated discriminators -> optim_g, two fused AdamW launches.  CUDA-event timing.
Imported by bench.py (key "vqvae_step"); stand-alone:   python tools/vqvae_step_bench.py [B] [iters] [--cpu]
`--cpu` / with_cpu: the REAL reference's step (SynthesizerTrn + MultiPeriodDiscriminator + the trainer's losses and AdamW pair, aug = identity)
on the host cores at a bounded batch, when a reference tree is on the box (baseline/_ref)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "baseline"))

SEG = 32            # segment_size 20480 / hop 640 (ttts/vqvae/config.json)
TEXT = 40
FLOP_PER_CLIP = 3.87e11     # FlopCounterMode over the reference's full step per sample (aug = identity)


def run(B=64, iters=3, with_cpu=False, cpu_batch=2):
    import torch
    import make_golden as MG
    from ttts_b200 import _lib as L
    from ttts_b200.vqvae.mel import spectrogram_torch
    from ttts_b200.vqvae.train_encoder import CudaKernels
    from ttts_b200.vqvae.train_step import TrainStep

    dev = torch.device("cuda")
    G, D = MG.step_params()
    G = {k: v.to(dev) for k, v in G.items()}
    D = {k: v.to(dev) for k, v in D.items()}
    g = torch.Generator(device="cuda").manual_seed(1234)
    wav = torch.clamp(0.1 * torch.randn(B, 23040, device=dev, generator=g), -1, 1)
    lengths = torch.full((B,), 36, dtype=torch.int64, device=dev)
    text = torch.randint(0, 256, (B, TEXT), device=dev, generator=g)
    text_lengths = torch.full((B,), TEXT, dtype=torch.int64, device=dev)
    E = torch.randn(1024, 192, device=dev, generator=g)
    eps_p, eps_q = torch.randn(B, 192, 36, device=dev, generator=g), torch.randn(B, 192, 36, device=dev, generator=g)
    ids = [int(i) % 5 for i in range(B)]
    lib = L.lib()
    lib.ttts_launch_count.restype = __import__("ctypes").c_ulonglong
    ts = TrainStep(CudaKernels(), G, D)

    def step():
        spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
        return ts.step(wav, spec, lengths, text, text_lengths, E, eps_p, eps_q, ids, SEG)
    out = step()
    torch.cuda.synchronize()
    l0 = lib.ttts_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        out = step()
    e1.record()
    enqueue = (time.perf_counter() - t0) / iters * 1e3              # host time to ENQUEUE a step (no synchronisation inside the loop)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters * 1e3
    ms = e0.elapsed_time(e1) / iters
    launches = (lib.ttts_launch_count() - l0) // iters
    res = {"workload": "VQ-VAE-GAN train step (BASELINE config 4): enc + VQ + dec + disc, generator + discriminator AdamW, segment 20480, aug = identity",
           "batch": B, "samples_per_clip": 23040, "ms_per_step": ms, "host_ms_per_step": wall, "host_enqueue_ms_per_step": enqueue, "msamples_per_s": B * 23040 / ms / 1e3,
           "gpu_launches_per_step": int(launches), "dtype": "f32 (wide convolutions: split-bf16 tcgen05, 3 bf16 products per fp32 product)",
           "step_tflops": FLOP_PER_CLIP * B / (ms * 1e-3) / 1e12,
           "flop_model": "3.87e11 FLOP per clip and step (SURVEY.md section 6, FlopCounterMode over the reference's step)",
           "losses": {k: float(v) for k, v in out.items()}}
    # roofline of the step's kernel families: one more step per family with a CUDA-event pair around each of its launches (host_util.cu:
    # 1 = the tcgen05 GEMM that the wide convolutions run on with split-bf16 operands, 3 = conv1d_wgrad2, 4 = the fp32 forward kernels, which
    # also compute the input gradients).  The GEMM family is measured against the measured dense bf16 peak (its FLOP count is the bf16 work:
    # three bf16 products per fp32 product), the exact-fp32 CUDA-core families against the fp32 FMA pipe (148 SMs x 128 lanes x 2 FLOP x max clock).
    import ctypes
    fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    try:
        bf16_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
        peak_src = "measured (MEASURED_PEAKS.json, burst)"
    except (OSError, KeyError):
        bf16_peak, peak_src = 1590.0, "fallback (B200_PROFILING.md)"
    roof = {}
    for kind, name, tensor in ((1, "gemm_bf16 / gemm2_bf16_kernel (tcgen05; split-bf16 convolutions)", True), (3, "conv1d_wgrad2_kernel", False),
                               (4, "conv1d_igemm_pipe_kernel / conv1d_direct_kernel (fp32 forward family, incl. input gradients)", False)):
        lib.ttts_prof_gemm_enable(kind)
        step()
        ms_k, fl_k, n_k = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.ttts_prof_gemm_read(ctypes.byref(ms_k), ctypes.byref(fl_k), ctypes.byref(n_k))
        lib.ttts_prof_gemm_enable(0)
        if ms_k.value > 0:
            tf = fl_k.value / (ms_k.value * 1e-3) / 1e12
            pk = bf16_peak if tensor else fma_peak
            roof[name] = {"launches": int(n_k.value), "ms_per_step": ms_k.value, "share_of_step": ms_k.value / ms, "tflops": tf,
                          "bound": "tensor (dense bf16, %s)" % peak_src if tensor else "fp32 FMA pipe", "peak_tflops": pk, "frac": tf / pk}
    if roof:
        top = max(roof, key=lambda k: roof[k]["ms_per_step"])
        res["roofline"] = {"bound": "tensor" if roof[top]["bound"].startswith("tensor") else "fp32 FMA pipe (exact-fp32 CUDA-core implicit GEMM)",
                           "kernel": top, "achieved": roof[top]["tflops"], "peak": roof[top]["peak_tflops"], "unit": "TFLOP/s", "frac": roof[top]["frac"],
                           "traffic": None, "kernels": roof}
    if with_cpu:
        try:
            res["cpu_baseline"] = cpu_reference_step(cpu_batch)
        except Exception as e:
            res["cpu_baseline"] = {"error": repr(e)[:300]}
    return res


def cpu_reference_step(B=2, steps=1):
    """the REAL reference's optimisation step on the host cores (the body of ttts/vqvae/train.py:330-406 with aug = identity, fp32, as
    tests/golden/make_golden.py::vqvae_full_step_case runs it), random-init modules, segment 32 frames, one warm-up + `steps` timed"""
    import torch
    import ref_loader
    gm = ref_loader.import_reference()
    if gm is None:
        return {"unavailable": "no reference tree on this box (baseline/_ref, /root/reference)"}
    from ttts.vqvae.vq2 import SynthesizerTrn, MultiPeriodDiscriminator
    from ttts.vqvae import losses as RL
    from ttts.utils import commons
    from ttts.utils.data_utils import spectrogram_torch, spec_to_mel_torch, mel_spectrogram_torch
    cfg = json.load(open(os.path.join(ref_loader.find_reference(), "ttts", "vqvae", "config.json")))
    torch.manual_seed(0)
    net_g = SynthesizerTrn(1025, SEG, **cfg["vqvae"]).train()
    net_d = MultiPeriodDiscriminator(False).train()
    cb = net_g.quantizer.vq.layers[0]._codebook
    cb.embed.copy_(torch.randn(1024, 192)); cb.embed_avg.copy_(cb.embed); cb.cluster_size.fill_(10); cb.inited.fill_(1)
    optim_g = torch.optim.AdamW(net_g.parameters(), 1e-4, betas=(0.8, 0.99), eps=1e-9)
    optim_d = torch.optim.AdamW(net_d.parameters(), 1e-4, betas=(0.8, 0.99), eps=1e-9)
    wav = torch.clamp(0.1 * torch.randn(B, 23040), -1, 1)
    lengths = torch.full((B,), 36, dtype=torch.int64)
    text = torch.randint(0, 256, (B, TEXT))
    text_lengths = torch.full((B,), TEXT, dtype=torch.int64)
    times = []
    for i in range(1 + steps):
        t0 = time.perf_counter()
        spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
        y_hat, kl_ssl, ids_slice, z_mask, (z, z_p, m_p, logs_p, m_q, logs_q), _ = net_g(wav, wav, lengths * 640, spec, spec, lengths, text.clone(), text_lengths)
        mel = spec_to_mel_torch(spec, 2048, 128, 32000, 0.0, None)
        y_mel = commons.slice_segments(mel, ids_slice, SEG)
        y_hat_mel = mel_spectrogram_torch(y_hat.squeeze(1), 2048, 128, 32000, 640, 2048, 0.0, None)
        y = commons.slice_segments(wav.unsqueeze(1), ids_slice * 640, SEG * 640)
        y_d_hat_r, y_d_hat_g, _, _ = net_d(y, y_hat.detach())
        loss_disc, _, _ = RL.discriminator_loss(y_d_hat_r, y_d_hat_g)
        optim_d.zero_grad(); loss_disc.backward(); optim_d.step()
        y_d_hat_r, y_d_hat_g, fmap_r, fmap_g = net_d(y, y_hat)
        loss_mel = torch.nn.functional.l1_loss(y_mel, y_hat_mel) * 45
        loss_kl = RL.kl_loss(z_p, logs_q, m_p, logs_p, z_mask) * 1.0
        loss_fm = RL.feature_loss(fmap_r, fmap_g)
        loss_gen, _ = RL.generator_loss(y_d_hat_g)
        total = loss_gen + loss_fm + loss_mel + kl_ssl * 1 + loss_kl
        optim_g.zero_grad(); total.backward(); optim_g.step()
        if i > 0:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return {"value": B * 23040 / t / 1e6, "unit": "Msamples/s", "s_per_step": t, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": "the REAL SynthesizerTrn + MultiPeriodDiscriminator step (fp32, aug = identity) at batch %d, %d timed step(s) after 1 warm-up" % (B, steps)}


if __name__ == "__main__":
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    print(json.dumps(run(int(a[0]) if a else 8, int(a[1]) if len(a) > 1 else 3, with_cpu="--cpu" in sys.argv)))
