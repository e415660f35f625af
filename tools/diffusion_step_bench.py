"""BASELINE.json config 5: one diffusion mel-refiner train step (AA_diffusion, 46 M parameters, under SpacedDiffusion.training_losses;
ttts/diffusion/train.py:156-203) at batch B x 1024 mel frames (latent 256 positions, reference clip 200 frames) through
`ttts_b200.diffusion.train_step.DiffusionStep.step` -- q_sample -> model -> MSE + VB loss -> backward -> grad-norm, clip 1.0, fused AdamW.
CUDA-event timing.  Imported by bench.py (key "diffusion_step"); stand-alone:   python tools/diffusion_step_bench.py [B] [iters] [--cpu]
`--cpu` / with_cpu: the REAL reference's micro-step + clip + AdamW on the host cores at a bounded batch, when a reference tree is on the box."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

T_MEL, T_LAT, T_REF = 1024, 256, 200


def flop_model(cfg, T=T_MEL, TL=T_LAT, TR=T_REF):
    """forward multiply-add FLOP of one sample (convolutions 2 Cin Cout K T, attention 4 T^2 C); a train step is 3x (dgrad + wgrad)"""
    C, L = cfg["model_channels"], cfg["num_layers"]
    conv = lambda cin, cout, k, t: 2.0 * cin * cout * k * t
    attn = lambda t: conv(C, 3 * C, 1, t) + 4.0 * t * t * C + conv(C, C, 1, t)
    res = lambda t: conv(C, C, 1, t) + conv(C, C, 3, t) + 2.0 * C * 2 * C
    f = conv(cfg["in_latent_channels"], C, 3, TL) + 3 * attn(TL)                        # latent_conditioner
    f += conv(cfg["in_channels"], C, 3, TR) + 3 * attn(TR)                              # refer_enc.0-3
    f += 4 * conv(C, C, 1, TR) + conv(C, C, 3, TR + 32) + 4 * attn(TR + 32)             # RefEncoder
    f += 3 * (res(T) + attn(T))                                                         # conditioning_timestep_integrator
    f += conv(cfg["in_channels"], C, 3, T) + conv(2 * C, C, 1, T)                       # inp_block, integrating_conv
    f += L * (res(T) + attn(T)) + 3 * res(T) + conv(C, cfg["out_channels"], 3, T)
    return f


def run(B=32, iters=3, with_cpu=False, cpu_batch=2):
    import ctypes
    import torch
    from ttts_b200 import _lib as L
    from ttts_b200.diffusion.kernels import DiffusionCudaKernels
    from ttts_b200.diffusion.params import default_config, init_params
    from ttts_b200.diffusion.train_step import DiffusionStep, normalize_tacotron_mel

    dev = torch.device("cuda")
    cfg = default_config()
    params = init_params(cfg, seed=0, device=dev, zero_proj_out=False)
    n_params = sum(v.numel() for v in params.values())
    g = torch.Generator(device="cuda").manual_seed(1234)
    mel = 2.0 * torch.randn(B, 100, T_MEL, device=dev, generator=g) - 5.0
    refer = 2.0 * torch.randn(B, 100, T_REF, device=dev, generator=g) - 5.0
    latent = torch.randn(B, cfg["in_latent_channels"], T_LAT, device=dev, generator=g)
    lib = L.lib()
    lib.ttts_launch_count.restype = ctypes.c_ulonglong
    ds = DiffusionStep(DiffusionCudaKernels(), params, cfg, seed=0)

    def step():
        batch = dict(x_start=normalize_tacotron_mel(mel), refer=normalize_tacotron_mel(refer), latent=latent,
                     noise=torch.randn(B, 100, T_MEL, device=dev, generator=g))
        return ds.step([batch])
    out = step()
    torch.cuda.synchronize()
    l0 = lib.ttts_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters * 1e3
    ms = e0.elapsed_time(e1) / iters
    launches = (lib.ttts_launch_count() - l0) // iters
    fl = 3.0 * flop_model(cfg) * B
    res = {"workload": "diffusion mel-refiner train step (BASELINE config 5): AA_diffusion 6 layers / 512 channels / 16 heads under SpacedDiffusion.training_losses "
                       "(epsilon, learned-range variance, MSE + VB), layer drop 0.1, unconditioned 0.1, clip 1.0 + AdamW",
           "batch": B, "mel_frames": T_MEL, "latent_positions": T_LAT, "refer_frames": T_REF, "parameters": int(n_params), "ms_per_step": ms,
           "host_ms_per_step": wall, "frames_per_s": B * T_MEL / ms * 1e3, "gpu_launches_per_step": int(launches), "dtype": "f32",
           "step_tflops": fl / (ms * 1e-3) / 1e12,
           "flop_model": "3 x forward FLOP (2 Cin Cout K T per convolution, 4 T^2 C per attention), no layer dropped: %.3e per sample and step" % (3.0 * flop_model(cfg)),
           "loss": float(out["loss"]), "grad_norm": float(out["grad_norm"])}
    fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    tc = os.environ.get("TTTS_DIFF_TC", "1") != "0"
    res["dtype"] = "f32 (wide convolutions: split-bf16 tcgen05, 3 bf16 products per fp32 product; attention / norms / loss: fp32 CUDA cores)" if tc else "f32"
    res["conv_path"] = "ttts_gemm_bf16 (tcgen05) with split-bf16 operands" if tc else "fp32 implicit GEMM on CUDA cores"
    res["roofline"] = {"bound": "fp32 FMA pipe (attention, the dominant kernels after the convolutions moved to tcgen05)" if tc else "fp32 FMA pipe (exact-fp32 CUDA-core kernels)",
                       "achieved": res["step_tflops"], "peak": fma_peak, "unit": "TFLOP/s", "frac": res["step_tflops"] / fma_peak, "traffic": None,
                       "note": "whole step (3 x forward FLOP, attention recompute not counted) against the fp32 FMA peak; with the convolutions on the "
                               "tensor cores the fraction can exceed what CUDA cores alone could reach"}
    if with_cpu:
        try:
            res["cpu_baseline"] = cpu_reference_step(cpu_batch)
        except Exception as e:
            res["cpu_baseline"] = {"error": repr(e)[:300]}
    return res


def cpu_reference_step(B=2, steps=1):
    """the REAL reference's step on the host cores: AA_diffusion(**config.yaml) in train() mode under SpacedDiffusion.training_losses, the loop
    body of ttts/diffusion/train.py:168-196 (accelerate absent: plain backward), fp32, one warm-up + `steps` timed"""
    import types
    import torch
    import ref_loader
    gm = ref_loader.import_reference()
    if gm is None:
        return {"unavailable": "no reference tree on this box (baseline/_ref, /root/reference)"}
    kd = types.ModuleType("k_diffusion"); ks = types.ModuleType("k_diffusion.sampling")
    ks.sample_dpmpp_2m = ks.sample_euler_ancestral = None; kd.sampling = ks
    sys.modules.setdefault("k_diffusion", kd); sys.modules.setdefault("k_diffusion.sampling", ks)
    from ttts.diffusion.aa_model import AA_diffusion, normalize_tacotron_mel
    from ttts.utils.diffusion import SpacedDiffusion, space_timesteps, get_named_beta_schedule
    torch.manual_seed(0)
    net = AA_diffusion(in_channels=100, out_channels=200, model_channels=512, num_heads=16, num_layers=6, in_latent_channels=512, dropout=0,
                       layer_drop=0.1).train()
    diffuser = SpacedDiffusion(use_timesteps=space_timesteps(1000, [1000]), model_mean_type="epsilon", model_var_type="learned_range", loss_type="mse",
                               betas=get_named_beta_schedule("linear", 1000), conditioning_free=False, conditioning_free_k=2.0)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01)
    mel = 2.0 * torch.randn(B, 100, T_MEL) - 5.0
    refer = 2.0 * torch.randn(B, 100, T_REF) - 5.0
    latent = torch.randn(B, 512, T_LAT)
    times = []
    for it in range(steps + 1):
        t0 = time.perf_counter()
        x_start, rf = normalize_tacotron_mel(mel), normalize_tacotron_mel(refer)
        t = torch.randint(0, 1000, (B,)).long()
        loss = diffuser.training_losses(model=net, x_start=x_start, t=t, model_kwargs={"latent": latent, "refer": rf})["loss"].mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step(); opt.zero_grad()
        times.append(time.perf_counter() - t0)
    s = min(times[1:])
    return {"value": B * T_MEL / s, "unit": "mel frames/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": "the REAL AA_diffusion + SpacedDiffusion.training_losses + clip + AdamW, batch %d x %d frames, fp32, %d step(s) after one warm-up" % (B, T_MEL, steps),
            "s_per_step": s, "loss": float(loss)}


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if args else 32
    iters = int(args[1]) if len(args) > 1 else 3
    if "--cpu-only" in sys.argv:
        print(json.dumps(cpu_reference_step(B)))
    else:
        print(json.dumps(run(B, iters, with_cpu="--cpu" in sys.argv)))
