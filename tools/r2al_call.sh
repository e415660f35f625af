# Round 2, GPU call AL: is the VQ-VAE-GAN step host-bound?  Host enqueue time per step, and a cProfile of the Python side of two steps.
mkdir -p gpurun_out
T=r2al
timeout 300 python - <<'PY' > gpurun_out/r2al_vqvae_host.txt 2>&1
import cProfile, pstats, io, json, sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import vqvae_step_bench as VB
r = VB.run(B=64, iters=2)
print(json.dumps({k: r[k] for k in ("ms_per_step", "host_ms_per_step", "host_enqueue_ms_per_step", "gpu_launches_per_step")}))
pr = cProfile.Profile(); pr.enable(); VB.run(B=64, iters=2); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
PY
head -c 7000 gpurun_out/r2al_vqvae_host.txt
