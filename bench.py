#!/usr/bin/env python
"""bench.py -- GPT-step audio-frames/sec of the UnifiedVoice train step (BASELINE.json north_star).

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a CUDA engine)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Workload (config.workload "cfg3"): UnifiedVoice 24L / d1024 / 16 heads, per-GPU batch 32, text 128, codes 1024 (T = 1156),
bf16 tensor-core operands with fp32 accumulate / residual / optimizer, dropout 0.1 active (training mode), synthetic tokens,
random-init weights.  A step = forward + backward + gradient all-reduce (N > 1) + global-norm clip + AdamW.
One "audio frame" = one VQ code position: frames/step = N * 32 * 1024.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same step driven through the
public `Trainer.train_step` with pinned HOST batches (H2D copies + loss D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg3": dict(layers=24, model_dim=1024, heads=16, B=32, TL=128, CL=1024),
    "cfg2": dict(layers=12, model_dim=512, heads=8, B=8, TL=128, CL=512),
    "tiny": dict(layers=2, model_dim=128, heads=2, B=2, TL=16, CL=32),        # tests/test_bench_contract_cpu.py only: not a bench configuration
}
GPT_KW = dict(max_text_tokens=800, max_mel_tokens=1600, number_text_tokens=256, start_text_token=255, number_mel_codes=1026,
              start_mel_token=1024, stop_mel_token=1025)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


def ncu_dram_bytes_per_launch():
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the newest committed
    `ncu --set full` capture of the shipped GEMM kernel (the QKV GEMM, M 36 992 x N 3 072 x K 1 024, whose algorithmic operand + output
    bytes are 309 MB; `python tools/gemm_prof.py` under ncu).  Measured under ncu, so it is read from the committed summary, never produced
    by this run."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[2-9]*_gemm_qkv_ncu_full.txt")), reverse=True)
    cands.append(os.path.join(ROOT, "profiles", "r1d_gemm2_pair_ncu_full.txt"))
    for path in cands:
        try:
            rd = wr = None
            for line in open(path):
                if line.startswith("dram__bytes_read.sum [Mbyte]"):
                    rd = float(line.split("]")[1].split("|")[0])
                if line.startswith("dram__bytes_write.sum [Mbyte]"):
                    wr = float(line.split("]")[1].split("|")[0])
            if rd is not None and wr is not None:
                return (rd + wr) * 1e6, "profiles/%s (QKV GEMM launch: 309 MB algorithmic)" % os.path.basename(path)
        except OSError:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def pick_cpu_threads(wl):
    """torch's intra-op pool with every hardware thread is often SLOWER on many-core hosts for these shapes; take the best of
    {all, 1/2, 1/4} threads on a 2-layer slice of the workload (a few seconds)."""
    import torch
    from oracle import gpt_oracle as O
    n = os.cpu_count() or 1
    cfg = O.default_config(layers=2, model_dim=wl["model_dim"], heads=wl["heads"])
    params = O.init_params(cfg, seed=0)
    batch = O.synthetic_batch(1, wl["TL"], wl["CL"], seed=1)
    best, best_t = n, None
    for th in sorted({n, max(1, n // 2), max(1, n // 4)}, reverse=True):
        torch.set_num_threads(th)
        O.loss_and_grads(params, cfg, *batch)
        t0 = time.perf_counter()
        O.loss_and_grads(params, cfg, *batch)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
    return best


def cpu_oracle_step_time(wl, B, steps, warm, threads):
    """The reference's CPU path (oracle port, fp32): fwd + bwd + clip + AdamW on a bounded sample (batch B of the workload)."""
    import torch
    from oracle import gpt_oracle as O
    torch.set_num_threads(threads)
    cfg = O.default_config(layers=wl["layers"], model_dim=wl["model_dim"], heads=wl["heads"])
    params = O.init_params(cfg, seed=0)
    state = {}
    text, tl, codes, wlens = O.synthetic_batch(B, wl["TL"], wl["CL"], seed=1234)
    times = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        _, _, _, grads = O.loss_and_grads(params, cfg, text, tl, codes, wlens)
        O.clip_and_adamw(params, grads, state, O.warmup_lr(i), i + 1)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return sum(times) / len(times)


def host_info():
    import torch
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    try:
        import transformers
        tv = transformers.__version__
    except Exception:
        tv = None
    return {"os_cpu_count": os.cpu_count(), "cpu_model": model, "torch": torch.__version__, "transformers": tv}


def reference_step_times(gm, wl, Bs, steps, warm, threads, checkpointing, autocast_bf16):
    """The reference's own train step on the host cores: the REAL ttts.gpt.model.UnifiedVoice driven by the loop body of
    ttts/gpt/train.py:99-121 restated line by line (accelerate is not installed: `accelerator.autocast()` = torch.autocast('cpu', bf16) or
    nothing, `accelerator.backward` = loss.backward(), `accelerator.clip_grad_norm_` = torch's) -- forward, 0.01 * loss_text + loss_mel,
    .item(), backward, get_grad_norm (the per-tensor .item() loop, train.py:22-31), clip 1.0, AdamW(lr 1e-4, betas (0.9, 0.96), wd 0.01),
    zero_grad, LambdaLR warm-up.  Training mode: the four GPT-2 dropouts are on and gradient checkpointing (the constructor default,
    model.py:256,297) recomputes every block."""
    import torch
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    kw = dict(GPT_KW, layers=wl["layers"], model_dim=wl["model_dim"], heads=wl["heads"])
    model = gm.UnifiedVoice(**kw, checkpointing=checkpointing).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.96), weight_decay=0.01)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda st: float(st / 500) if st < 500 else 1)
    from ttts_b200.gpt import synth
    text, tl, codes, wlens = synth.synthetic_batch(Bs, wl["TL"], wl["CL"], seed=1234)
    times = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        if autocast_bf16:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                loss_text, loss_mel, _ = model(text, tl, codes.clone(), wlens)
                loss = 0.01 * loss_text + 1.0 * loss_mel
        else:
            loss_text, loss_mel, _ = model(text, tl, codes.clone(), wlens)
            loss = 0.01 * loss_text + 1.0 * loss_mel
        total = loss.item()
        loss.backward()
        total_norm = 0.0
        for p_ in model.parameters():                       # get_grad_norm, train.py:22-31
            if p_.grad is not None:
                total_norm += p_.grad.data.norm(2).item() ** 2
        total_norm = total_norm ** 0.5
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step(); opt.zero_grad(); sched.step()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    assert total == total and total_norm > 0
    return sum(times) / len(times), total


def run_reference(args, wl):
    """bench.py --impl reference: the reference's CPU implementation of the path on the box's host cores.  The REAL reference modules when a
    reference tree is on the box (baseline/_ref, installed by baseline/install_ref.py, or /root/reference): kind "reference"; otherwise the
    oracle port (unit-tested identical to it): kind "port".  BASELINE.md section 3: fp32 with gradient checkpointing (the reference's default
    configuration), fp32 without, bf16 autocast with and without, on a bounded sample (batch 2 of the workload's shape; frames/s does not depend on the
    batch to first order, so the cfg3 batch of 32 is this x16: flagged `extrapolated`).  `value` = the FASTEST variant, so the ratio the driver
    forms against it is the conservative one."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    gm = None
    why_port = None
    try:
        import ref_loader
        gm = ref_loader.import_reference()
        if gm is None:
            why_port = "no reference tree on this box (baseline/_ref, /root/reference)"
    except Exception as e:                                   # the port always exists
        why_port = "reference import failed: " + repr(e)[:200]
    threads = pick_cpu_threads(wl)
    info = host_info()
    info["threads_used"] = threads
    info["threads_note"] = "best of {all, 1/2, 1/4} hardware threads on a 2-layer probe (torch's intra-op pool with every thread is slower on many-core hosts)"
    variants = {}
    if gm is not None:
        Bs = 2
        budget = 150.0
        plan = [("fp32_gradckpt", True, False), ("fp32", False, False), ("bf16_autocast", False, True), ("bf16_autocast_gradckpt", True, True)]
        t_probe, _ = reference_step_times(gm, wl, Bs, 1, 0, threads, False, False)
        steps = max(1, min(args.steps, int(budget / (len(plan) * 1.4 * max(t_probe, 1e-3))) - 1))
        for name, ckpt, ac in plan:
            t, loss = reference_step_times(gm, wl, Bs, steps, 1, threads, ckpt, ac)
            variants[name] = {"s_per_step": t, "frames_per_s": Bs * wl["CL"] / t, "loss": loss}
        best = max(variants, key=lambda k: variants[k]["frames_per_s"])
        t = variants[best]["s_per_step"]
        kind = "reference"
        sample = ("the REAL ttts.gpt.model.UnifiedVoice (%s) driven by the loop body of ttts/gpt/train.py:99-121, training mode, batch %d of the %s shape, "
                  "%d timed steps after 1 warm-up per variant; value = fastest variant (%s); the reference's default configuration is fp32_gradckpt"
                  % (ref_loader.find_reference(), Bs, args.workload, steps, best))
        dtype = "bf16" if "bf16" in best else "f32"
    else:
        Bs = 1
        t1 = cpu_oracle_step_time(wl, Bs, 1, 0, threads)
        steps = max(1, min(args.steps, int(150.0 / max(t1, 1e-3))))
        t = cpu_oracle_step_time(wl, Bs, steps, 0, threads) if steps > 1 else t1
        kind, best, dtype = "port", "fp32_port", "f32"
        sample = "oracle port (fp32, no grad-ckpt) of the reference step at batch %d of the %s shape, %d timed steps (%s)" % (Bs, args.workload, steps, why_port)
    args.steps = steps
    fps = Bs * wl["CL"] / t
    out = {
        "impl": "reference", "metric": "gpt_step_audio_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
        "data": "synthetic", "config": {"workload": args.workload, "model": "UnifiedVoice %dL/d%d/H%d" % (wl["layers"], wl["model_dim"], wl["heads"]),
                                       "per_gpu_batch": wl["B"], "text_len": wl["TL"], "code_len": wl["CL"], "seq_len": wl["TL"] + wl["CL"] + 4,
                                       "sample_batch": Bs, "extrapolated": Bs != wl["B"], "dropout": 0.1, "step": "fwd+bwd+clip+AdamW"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample, "variant": best, "variants": variants,
                         "host": info},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-vq-encode", action="store_true")
    ap.add_argument("--no-vqvae-step", action="store_true")
    ap.add_argument("--no-diffusion-step", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the headline): the workload's batch PER GPU; strong: that batch split over the GPUs (SURVEY.md 8d secondary line)")
    ap.add_argument("--profile-run", action="store_true", help="for ncu runs only: honour --warmup < 3 (numbers printed are not bench values)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    from ttts_b200 import _lib as L
    from ttts_b200.gpt.model import UnifiedVoice
    from ttts_b200.gpt.train import FusedStep, Trainer
    from ttts_b200.gpt import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = args.warmup if args.profile_run else max(args.warmup, 3)

    cfg_json = {"train": {"train_steps": 10 ** 9, "val_freq": 10 ** 9, "save_freq": 10 ** 9, "keep_ckpts": 0, "lr": 1e-4, "logs_folder": "/tmp/ttts_b200_logs",
                          "text_weight": 0.01, "mel_weight": 1, "accumulate_num": 1},
                "gpt": dict(GPT_KW, layers=wl["layers"], model_dim=wl["model_dim"], heads=wl["heads"])}
    B, TL, CL = wl["B"], wl["TL"], wl["CL"]
    if args.scaling == "strong":
        assert B % world == 0, "strong scaling: global batch %d is not divisible by %d GPUs" % (B, world)
        B //= world
    torch.manual_seed(0)

    # host batches (pinned) -- a few distinct ones, rank-dependent
    def host_batch(i):
        text, tl, codes, wlens = synth.synthetic_batch(B, TL, CL, seed=1234 + 1000 * rank + i)
        return {"padded_text": text.pin_memory(), "text_lengths": tl.pin_memory(), "padded_qmel": codes.pin_memory(), "wav_lens": wlens.pin_memory()}
    batches = [host_batch(i) for i in range(4)]
    trainer = Trainer(cfg=cfg_json, dataloader=[batches[0]], device=dev, logs=False)
    model = trainer.gpt
    model.train()
    model.dropout_p = args.dropout
    fused = trainer.fused
    dev_batches = [[b[k].to(dev) for k in ("padded_text", "text_lengths", "padded_qmel", "wav_lens")] for b in batches]
    lib = L.lib()
    lib.ttts_launch_count.restype = __import__("ctypes").c_ulonglong

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timed region ----------------
    for i in range(warm):
        fused(*dev_batches[i % 4], clip_inputs=False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    import ctypes
    l0 = lib.ttts_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    step_ev = []
    for i in range(args.steps):
        fused(*dev_batches[i % 4], clip_inputs=False)
        ev = torch.cuda.Event(enable_timing=True); ev.record(); step_ev.append(ev)      # per-step marks (no sync): p10 / median / p90
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    marks = [e0] + step_ev
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
    pct = lambda q: per_step[min(len(per_step) - 1, int(q * len(per_step)))]
    l1 = lib.ttts_launch_count()
    # roofline pass: the SAME K steps once more with a CUDA-event pair around every GEMM launch (on the launch stream).  Kept out of the
    # headline region above: 588 event records per step serialise neighbouring kernels and cost ~2 % of the step (r1k: 101.0 vs 98.9 ms).
    lib.ttts_prof_gemm_enable(1)
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ep0.record()
    for i in range(args.steps):
        fused(*dev_batches[i % 4], clip_inputs=False)
    ep1.record()
    barrier()
    ms_prof_step = ep0.elapsed_time(ep1) / args.steps
    gms, gfl, gn = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    lib.ttts_prof_gemm_read(ctypes.byref(gms), ctypes.byref(gfl), ctypes.byref(gn))
    lib.ttts_prof_gemm_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    frames_step = world * B * CL
    value = frames_step / (ms_step * 1e-3)

    # ---------------- end-to-end through the public Trainer API, host batches ----------------
    e2e = None
    if not args.no_e2e:
        for i in range(2):
            trainer.train_step(batches[i % 4])
        barrier()
        e0.record()
        for i in range(args.steps):
            trainer.train_step(batches[i % 4])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item() / args.steps
        h2d = sum(v.numel() * v.element_size() for v in batches[0].values())
        e2e = {"value": frames_step / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    fl_step = synth.flops_per_step(wl["layers"], wl["model_dim"], B, TL, CL)     # per GPU, algorithmic
    gemm_tflops = (gfl.value / (gms.value * 1e-3)) / 1e12 if gms.value > 0 else None
    traffic, traffic_src = ncu_dram_bytes_per_launch()
    roofline = {
        "bound": "tensor", "kernel": "gemm2_bf16_kernel<A_MN,B_MN,PAIR> (tcgen05.mma cta_group::2, TMA, TMEM; all GEMMs of the step)",
        "achieved": gemm_tflops, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
        "frac": (gemm_tflops / pk["bf16_sustained"]) if gemm_tflops else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": pk["src"] + " (sustained bf16 cuBLAS, MEASURED_PEAKS.json)",
        "launches": int(gn.value), "kernel_ms_per_step": gms.value / args.steps,
        "kernel_share_of_step": (gms.value / args.steps) / ms_prof_step,
        "measured_over": "%d further steps identical to the timed ones, one CUDA-event pair per GEMM launch on its stream (%.2f ms/step with the events; the headline region runs without them)" % (args.steps, ms_prof_step),
        "step_tflops": fl_step / (ms_step * 1e-3) / 1e12, "step_frac": fl_step / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"],
    }
    out = {
        "metric": "gpt_step_audio_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "step_ms_rank0": {"p10": pct(0.1), "median": pct(0.5), "p90": pct(0.9)}, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "model": "UnifiedVoice %dL/d%d/H%d" % (wl["layers"], wl["model_dim"], wl["heads"]), "per_gpu_batch": B,
                   "global_batch": B * world, "text_len": TL, "code_len": CL, "seq_len": TL + CL + 4, "parallelism": "dp%d" % world,
                   "dropout": args.dropout, "l2": "working set (~33 GB activations + 5 GB optimizer state per step) far exceeds the 126 MB L2; no flush needed",
                   "step": "fwd+bwd+allreduce+clip+AdamW"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(l1 - l0), "roofline": roofline,
    }
    if not args.no_cpu_baseline and world == 1:
        threads = pick_cpu_threads(wl)
        tcpu = cpu_oracle_step_time(wl, 1, 1, 0, threads)
        out["cpu_baseline"] = {"value": CL / tcpu, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": "1 step (fwd+bwd+clip+AdamW, fp32, no grad-ckpt) of the oracle port at batch 1 of the %s shape" % args.workload}
    if world == 1 and not args.profile_run and args.workload == "cfg3":
        # secondary: BASELINE config 2 (12L/d512, batch 8, 512 codes) -- too small to fill the machine, reported for completeness
        try:
            del trainer, fused, model
            torch.cuda.empty_cache()
            w2 = WORKLOADS["cfg2"]
            cfg2 = {"train": cfg_json["train"], "gpt": dict(GPT_KW, layers=w2["layers"], model_dim=w2["model_dim"], heads=w2["heads"])}
            b2 = [t.to(dev) for t in synth.synthetic_batch(w2["B"], w2["TL"], w2["CL"], seed=77)]
            tr2 = Trainer(cfg=cfg2, dataloader=[None], device=dev, logs=False)
            tr2.gpt.train(); tr2.gpt.dropout_p = args.dropout
            for _ in range(5):
                tr2.fused(*b2, clip_inputs=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                tr2.fused(*b2, clip_inputs=False)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / 20
            fl2 = synth.flops_per_step(w2["layers"], w2["model_dim"], w2["B"], w2["TL"], w2["CL"])
            out["cfg2"] = {"ms_per_step": ms2, "frames_per_s": w2["B"] * w2["CL"] / (ms2 * 1e-3), "step_tflops": fl2 / (ms2 * 1e-3) / 1e12,
                           "config": "UnifiedVoice 12L/d512/H8, batch 8, text 128, codes 512, 20 steps after 5 warm-up"}
            del tr2
        except Exception as e:
            out["cfg2"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_vq_encode:
        # second BASELINE.json metric: VQ-encode Msamples/s (stft -> ref_enc -> enc_p -> proj -> codes, B = 64 clips)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import vq_encode_bench
            out["vq_encode"] = vq_encode_bench.run(hbm_gbs=pk["hbm_gbs"], with_cpu=not args.no_cpu_baseline, bf16_tflops=pk["bf16_burst"])
        except Exception as e:            # the GPT line must still be printed
            out["vq_encode"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_vqvae_step and not args.profile_run and args.workload == "cfg3":
        # BASELINE config 4: one VQ-VAE-GAN train step (enc + VQ + dec + disc), batch 64 x 23 040 samples, through TrainStep.step
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import vqvae_step_bench
            torch.cuda.empty_cache()
            out["vqvae_step"] = vqvae_step_bench.run(B=64, iters=3, with_cpu=not args.no_cpu_baseline)
        except Exception as e:
            out["vqvae_step"] = {"error": repr(e)[:300]}
    if world == 1 and not args.no_diffusion_step and not args.profile_run and args.workload == "cfg3":
        # BASELINE config 5: one diffusion mel-refiner train step (AA_diffusion under training_losses), batch 32 x 1024 frames, through DiffusionStep.step
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import diffusion_step_bench
            torch.cuda.empty_cache()
            out["diffusion_step"] = diffusion_step_bench.run(B=32, iters=3, with_cpu=not args.no_cpu_baseline)
        except Exception as e:
            out["diffusion_step"] = {"error": repr(e)[:300]}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
