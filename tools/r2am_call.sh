# Round 2, GPU call AM (last of the round): raw-stream accessor + fused weight-operand kernel on the host-bound tapes: the whole GPU suite, smoke, A/B.
mkdir -p gpurun_out
T=r2am
S=$(date +%s)
timeout 150 python -m pytest tests -m gpu -q -x -rf > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "== pytest -m gpu rc=$? ($(( $(date +%s) - S )) s)"; tail -n 1 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest_gpu.log | cut -c1-300 | head -12
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  enqueue %.1f  launches %d" % (d["ms_per_step"], d.get("host_enqueue_ms_per_step", -1), d["gpu_launches_per_step"]), d.get("losses", d.get("loss")))'
timeout 100 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae.json 2> gpurun_out/${T}_vqvae.err; python -c "$P" gpurun_out/${T}_vqvae.json; tail -n 2 gpurun_out/${T}_vqvae.err | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "== smoke rc=$?"; tail -n 1 gpurun_out/${T}_smoke.log | cut -c1-300
TTTS_GEMM_WPREP=0 timeout 100 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_torchprep.json 2> gpurun_out/${T}_vqvae_torchprep.err; python -c "$P" gpurun_out/${T}_vqvae_torchprep.json
timeout 60 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion.json 2> gpurun_out/${T}_diffusion.err; python -c "$P" gpurun_out/${T}_diffusion.json
