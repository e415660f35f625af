"""Diagnosis of the training-tape gradients on the GPU: the decoder graph over the CUDA kernels against torch.autograd of the decoder oracle
run ON THE GPU in true fp32 (TF32 off), per tensor; run-to-run determinism; and the failing entries of the golden comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import decoder_oracle as DO
from ttts_b200.vqvae.train_decoder import DecoderGraph
from ttts_b200.vqvae.train_encoder import CudaKernels

dec = np.load(os.path.join(ROOT, "tests", "golden", "decoder.npz"))
P = {k: v.cuda() for k, v in DO.init_params(seed=9).items()}
z, g = torch.tensor(dec["z"]).cuda(), torch.tensor(dec["g"]).cuda()
R = torch.randn(dec["y"].shape, generator=torch.Generator().manual_seed(32)).cuda()
print("z", tuple(z.shape), "g", tuple(g.shape), "y", dec["y"].shape)

def tape():
    graph = DecoderGraph(CudaKernels(), P)
    y = graph.forward(z, g)
    return y.v.clone(), {k: v.clone() for k, v in graph.backward(R).items()}
y1, g1 = tape(); y2, g2 = tape()
Pr = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
yo = DO.generator(Pr, z, g)
(yo * R).sum().backward()
print("forward rel vs gpu-torch oracle %.2e ; vs golden %.2e" % (float((y1 - yo).norm() / yo.norm()), float(np.linalg.norm(y1.cpu().numpy() - dec["y"]) / np.linalg.norm(dec["y"]))))
rows = []
for k in g1:
    go = Pr[k].grad
    rows.append((float((g1[k] - go).norm() / (go.norm() + 1e-30)), float((g1[k] - g2[k]).norm() / (go.norm() + 1e-30)), k, tuple(go.shape), float(go.norm())))
rows.sort(reverse=True)
print("worst tensors: rel-to-oracle, run-to-run, name, shape, |g|")
for r in rows[:14]: print("  %.3e  %.3e  %-50s %s %.3e" % r)
print("median rel %.2e" % rows[len(rows) // 2][0])
k = rows[0][2]
d = (g1[k] - Pr[k].grad).flatten(); i = int(d.abs().argmax())
print("worst tensor", k, "largest element error at", i, float(d[i]), "of value", float(Pr[k].grad.flatten()[i]), "n_bad(>1e-3|g|max)", int((d.abs() > 1e-3 * Pr[k].grad.abs().max()).sum()), "of", d.numel())
names = [str(n) for n in dec["names"]]
bad = []
for i, k in enumerate(names):
    gk = g1[k].cpu(); dd = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i)); sc = float(dec["norm"][i])
    bad.append((abs(float((gk * dd).sum()) - float(dec["proj"][i])) / sc, abs(float(gk.norm()) - sc) / sc, k))
bad.sort(reverse=True)
print("golden comparison, worst (proj err / scale, norm err / scale):")
for b in bad[:8]: print("  %.3e %.3e %s" % b)
# the same with the oracle's gpu gradients in place of the tape's: is the GOLDEN comparison itself that noisy?
bad = []
for i, k in enumerate(names):
    gk = Pr[k].grad.cpu(); dd = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i)); sc = float(dec["norm"][i])
    bad.append((abs(float((gk * dd).sum()) - float(dec["proj"][i])) / sc, abs(float(gk.norm()) - sc) / sc, k))
bad.sort(reverse=True)
print("gpu-torch ORACLE vs golden, worst:")
for b in bad[:5]: print("  %.3e %.3e %s" % b)

# ---- generator step: which tensors fail, and by how much
import make_golden as MG
from ttts_b200.vqvae.mel import spectrogram_torch
from ttts_b200.vqvae.train_step import GeneratorStep
zz = np.load(os.path.join(ROOT, "tests", "golden", "vqvae_step.npz"))
G, D = MG.step_params()
wav, lengths, text, text_lengths, E = MG.step_inputs()
torch.manual_seed(0)
eps_p, eps_q = torch.randn(3, 192, 36), torch.randn(3, 192, 36)
ids = (torch.rand([3]) * (lengths - 8 + 1)).to(torch.long).tolist()
c = lambda t: t.cuda()
def gstep():
    step = GeneratorStep(CudaKernels(), {k: c(v) for k, v in G.items()}, {k: c(v) for k, v in D.items()})
    spec = spectrogram_torch(c(wav), 2048, 640, 2048, center=False)
    out = step.forward(c(wav), spec, c(lengths), c(text), c(text_lengths), c(E), c(eps_p), c(eps_q), ids, 8)
    return {k: float(out[k].v) for k in ("loss_gen", "loss_fm", "loss_mel", "kl_ssl", "loss_kl", "total")}, {k: v.clone() for k, v in step.backward().items()}
o1, s1 = gstep(); o2, s2 = gstep()
print("generator step losses", o1, "golden", {k: float(zz[k]) for k in o1})
names = [str(n) for n in zz["names"]]
bad = []
for i, k in enumerate(names):
    gk = s1[k].cpu(); dd = torch.randn(gk.shape, generator=torch.Generator().manual_seed(i)); sc = float(zz["norm"][i])
    rr = float((s1[k] - s2[k]).norm()) / (sc + 1e-30)
    bad.append((abs(float((gk * dd).sum()) - float(zz["proj"][i])) / (sc + 1e-30), abs(float(gk.norm()) - sc) / (sc + 1e-30), rr, k, sc))
bad.sort(reverse=True)
print("generator step vs golden, worst (proj err / scale, norm err / scale, run-to-run / scale, name, scale):")
for b in bad[:20]: print("  %.3e %.3e %.3e %s %.3e" % b)
