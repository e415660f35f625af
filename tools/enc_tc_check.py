"""Whole-encoder effect of the tensor-core convolutions: golden comparison (tests/golden/encoder.npz, minted from the REAL reference) and the
64-clip encode time, for the path selected by TTTS_CONV_TC (0 = exact fp32 CUDA-core kernels, 1 = split-bf16 tcgen05)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import encoder_oracle as EO
from oracle import vq_mel_oracle as V
from ttts_b200.vqvae.encoder import VQEncoder
enc = np.load(os.path.join(ROOT, "tests", "golden", "encoder.npz"))
m = VQEncoder(); m.load_state_dict(EO.init_params(seed=5), strict=False); m = m.cuda().eval()
cb = m.quantizer.vq.layers[0]._codebook; cb.embed.copy_(torch.tensor(enc["E"])); cb.inited.fill_(1)
out = m(torch.tensor(enc["wav"]).cuda(), lengths=torch.tensor(enc["lengths"]).cuda(), eps=torch.tensor(enc["eps"]).cuda())
rel = lambda a, b: float(np.linalg.norm(a.cpu().numpy().astype(np.float64) - b) / np.linalg.norm(b))
xn = np.ascontiguousarray(enc["x"].transpose(0, 2, 1)).reshape(-1, 192)
margin = V.vq_margin(xn, enc["E"], enc["codes"].reshape(-1))
flips = out["codes"].cpu().numpy().reshape(-1) != enc["codes"].reshape(-1)
print("TTTS_CONV_TC=%s  rel: ge %.2e m %.2e logs %.2e z %.2e x %.2e | code flips %d (largest margin among flips %.2e; smallest margin overall %.2e)" % (
    os.environ.get("TTTS_CONV_TC", "1"), rel(out["ge"], enc["ge"]), rel(out["m"], enc["m"]), rel(out["logs"], enc["logs"]), rel(out["z"], enc["z"]),
    rel(out["x"], enc["x"]), int(flips.sum()), float(margin[flips].max()) if flips.any() else 0.0, float(margin.min())), flush=True)
g = torch.Generator(device="cuda").manual_seed(1234)
wav = torch.clamp(0.1 * torch.randn(64, 23040, device="cuda", generator=g), -1, 1)
for _ in range(3): m(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): m(wav)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
ms_g = None
try:
    for _ in range(3): m.encode_graphed(wav)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): m.encode_graphed(wav)
    e1.record(); torch.cuda.synchronize(); ms_g = e0.elapsed_time(e1) / 10
except Exception as e:
    print("graphed encode failed:", repr(e)[:200])
print("encode 64 clips: %.3f ms = %.1f Msamples/s ; graphed %s" % (ms, 64 * 23040 / ms / 1e3, "%.3f ms = %.1f Msamples/s" % (ms_g, 64 * 23040 / ms_g / 1e3) if ms_g else "-"), flush=True)
