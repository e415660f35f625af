"""Autoregressive code generation rate: codes/s of `UnifiedVoice.inference_speech`-style decoding on the cfg3 model
(24L / d1024) for B sequences, uncached (the training forward over the sequence so far, what ttts/api_zh.py:51 runs) vs the KV-cache
decode step (csrc/gpt_decode.cu), eager launches and CUDA-graph replay.  Also reports the step's achieved HBM bandwidth against the
algorithmic bytes of one step: the bf16 parameter shadow of the layers + mel head (every weight byte is read once per step) plus the
cache rows read.  Prints one JSON object.   python tools/decode_bench.py [B] [prompt_codes] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ttts_b200.gpt import engine as E
    from ttts_b200.gpt.model import UnifiedVoice

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    m0 = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))                           # fallback: B200_PROFILING.md
    L_, d, H, TL = 24, 1024, 16, 128
    torch.manual_seed(0)
    model = UnifiedVoice(layers=L_, model_dim=d, heads=H, max_text_tokens=800, max_mel_tokens=1600, number_text_tokens=256, start_text_token=255,
                         number_mel_codes=1026, start_mel_token=1024, stop_mel_token=1025).cuda().eval()
    eng = model._engine()
    eng.refresh_shadow(force=True)
    g = torch.Generator(device="cuda").manual_seed(1234)
    text = torch.randint(1, 255, (B, TL), device="cuda", generator=g)
    codes = torch.randint(0, 1024, (B, m0 + steps + 8), device="cuda", generator=g)
    out = {"B": B, "prompt_codes": m0, "steps": steps, "model": "UnifiedVoice 24L/d1024/H16", "text_len": TL}

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # uncached: one training forward over [text ; codes so far] per code (sequence grows from m0 to m0 + steps)
    def uncached(i):
        n = m0 + i
        wav = torch.full((B,), (n + 1) * 1024, dtype=torch.int64, device="cuda")
        eng.forward(text, codes, wav, TL, n, save=False)
    uncached(0)
    n_unc = min(steps, 64)
    ms_unc = timed(uncached, n_unc)
    out["uncached_ms_per_code"] = ms_unc
    out["uncached_codes_per_s"] = B / ms_unc * 1e3

    # cached
    T_need = TL + 3 + m0 + steps + 8
    eng.decode_setup(B, T_need)
    wav = torch.full((B,), (m0 + 1) * 1024, dtype=torch.int64, device="cuda")

    def prefill():
        io = eng.forward(text, codes, wav, TL, m0, save=True)
        eng.kv_prefill(io, TL + 3 + m0)
    prefill()
    for graph in (False, True):
        prefill()
        eng.decode_step(codes, TL + 2, 0, graph=graph)                 # warm-up (and capture)
        ms = timed(lambda i: eng.decode_step(codes, TL + 2, 0, graph=graph), steps)
        key = "graph" if graph else "eager"
        out["cached_%s_ms_per_code" % key] = ms
        out["cached_%s_codes_per_s" % key] = B / ms * 1e3
        # algorithmic bytes of a step: layer weights (12 d^2 bf16) + mel head + the K / V rows read (average cache length over the run)
        t_avg = TL + 3 + m0 + steps / 2
        by = L_ * 12 * d * d * 2 + 1026 * d * 2 + L_ * 2 * B * H * t_avg * 64 * 2
        out["cached_%s_gbs" % key] = by / ms / 1e6
        out["cached_%s_frac_hbm" % key] = by / ms / 1e6 / hbm
    out["hbm_peak_gbs"] = hbm
    out["speedup_graph_vs_uncached"] = out["uncached_ms_per_code"] / out["cached_graph_ms_per_code"]
    # agreement: the last cached step's logits vs the uncached forward's logits at the same position
    n = m0 + steps + 1                                                   # after warm-up + `steps` steps the slot is TL + 3 + m0 + 1 + steps
    got = eng._dec["logits"].clone()
    wavn = torch.full((B,), (n + 1) * 1024, dtype=torch.int64, device="cuda")
    eng.forward(text, codes, wavn, TL, n, save=False)
    ld = E.L.lib().ttts_gpt_logits_ld(1026)
    ref = eng.ws_view(E.WS_MEL_LOGITS, B, TL, n, False, torch.bfloat16, (B, n + 2, ld))[:, n, :1026].float()
    out["cached_vs_uncached_logits_rel"] = float((got - ref).norm() / ref.norm())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
