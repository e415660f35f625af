# Round 2, GPU call AD: tap-concatenated GEMM convolutions (all taps of a dil = 1 layer in one reduction over an OVERLAPPED operand view; input
# gradient of stride-1 layers through the same route with flipped taps; weight gradient with the taps as column blocks): parity + A/B.
mkdir -p gpurun_out
T=r2ad
timeout 600 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -rf -k "tensor_core" > gpurun_out/${T}_pytest_conv.log 2>&1
echo "== pytest conv (tap-concat) rc=$?"; tail -1 gpurun_out/${T}_pytest_conv.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest_conv.log | cut -c1-250 | head -12
TTTS_GEMM_TAPCAT=0 timeout 600 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -rf -k "tensor_core" > gpurun_out/${T}_pytest_conv_pertap.log 2>&1
echo "== pytest conv (per tap) rc=$?"; tail -1 gpurun_out/${T}_pytest_conv_pertap.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest_conv_pertap.log | cut -c1-250 | head -12
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py tests/test_gpu_kernels.py -m gpu -q -rf -k "not tensor_core" > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -1 gpurun_out/${T}_pytest.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-250 | head -20
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  launches %d" % (d["ms_per_step"], d["gpu_launches_per_step"]), {k[:24]: (round(v["ms_per_step"],1), round(v["tflops"],1)) for k,v in d.get("roofline",{}).get("kernels",{}).items()}, d.get("losses", d.get("loss")))'
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_tapcat.json 2> gpurun_out/${T}_vqvae_tapcat.err; python -c "$P" gpurun_out/${T}_vqvae_tapcat.json; tail -n 3 gpurun_out/${T}_vqvae_tapcat.err | cut -c1-300
TTTS_GEMM_WCAT=0 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_nowcat.json 2> gpurun_out/${T}_vqvae_nowcat.err; python -c "$P" gpurun_out/${T}_vqvae_nowcat.json
TTTS_GEMM_TAPCAT=0 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_pertap.json 2> gpurun_out/${T}_vqvae_pertap.err; python -c "$P" gpurun_out/${T}_vqvae_pertap.json
timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion_tapcat.json 2> gpurun_out/${T}_diffusion_tapcat.err; python -c "$P" gpurun_out/${T}_diffusion_tapcat.json; tail -n 3 gpurun_out/${T}_diffusion_tapcat.err | cut -c1-300
TTTS_GEMM_TAPCAT=0 timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion_pertap.json 2> gpurun_out/${T}_diffusion_pertap.err; python -c "$P" gpurun_out/${T}_diffusion_pertap.json
