# Round 2, GPU call P: the diffusion step's wide convolutions on the tcgen05 GEMM with split-bf16 operands (fwd / dgrad / wgrad), grouped dgrad
# tap walk, sliced bias gradient; parity + A/B.
mkdir -p gpurun_out
T=r2p
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -30
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  TFLOP/s %.1f  launches %d" % (d["ms_per_step"], d["step_tflops"], d["gpu_launches_per_step"]), d.get("losses", d.get("loss")))'
timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion_tc.json 2> gpurun_out/${T}_diffusion_tc.err; python -c "$P" gpurun_out/${T}_diffusion_tc.json; tail -n 3 gpurun_out/${T}_diffusion_tc.err | cut -c1-300
TTTS_DIFF_TC=0 timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion_fp32.json 2> gpurun_out/${T}_diffusion_fp32.err; python -c "$P" gpurun_out/${T}_diffusion_fp32.json
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae.json 2> gpurun_out/${T}_vqvae.err; python -c "$P" gpurun_out/${T}_vqvae.json; tail -n 3 gpurun_out/${T}_vqvae.err | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -n 2 gpurun_out/${T}_smoke.log | cut -c1-400
