# Round 2, GPU call AE: narrow (32- / 64-channel) dil = 1 layers and the ConvTranspose1d phases / weight gradient on the tap-concatenated GEMM route.
mkdir -p gpurun_out
T=r2ae
timeout 600 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -rf -k "tensor_core or conv_transpose" > gpurun_out/${T}_pytest_conv.log 2>&1
echo "== pytest conv rc=$?"; tail -1 gpurun_out/${T}_pytest_conv.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest_conv.log | cut -c1-250 | head -20
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest encoder rc=$?"; tail -1 gpurun_out/${T}_pytest.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-250 | head -20
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  launches %d" % (d["ms_per_step"], d["gpu_launches_per_step"]), {k[:24]: (round(v["ms_per_step"],1), round(v["tflops"],1)) for k,v in d.get("roofline",{}).get("kernels",{}).items()}, d.get("losses", d.get("loss")))'
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_narrow.json 2> gpurun_out/${T}_vqvae_narrow.err; python -c "$P" gpurun_out/${T}_vqvae_narrow.json; tail -n 3 gpurun_out/${T}_vqvae_narrow.err | cut -c1-300
TTTS_GEMM_NARROW=0 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_wide_only.json 2> gpurun_out/${T}_vqvae_wide_only.err; python -c "$P" gpurun_out/${T}_vqvae_wide_only.json
