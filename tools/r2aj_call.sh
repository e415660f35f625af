# Round 2, GPU call AJ: dilated stride-1 layers as ordinary convolutions over de-interleaved sub-clips (tap-concatenated route): parity + A/B.
mkdir -p gpurun_out
T=r2aj
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -1 gpurun_out/${T}_pytest.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-250 | head -20
timeout 200 python tools/gemm_conv_prof.py 64 128 2560 128 11 1 3 15 | tee gpurun_out/${T}_gemm_conv_ab.txt
TTTS_GEMM_DEINT=0 timeout 200 python tools/gemm_conv_prof.py 64 128 2560 128 11 1 3 15 | tee -a gpurun_out/${T}_gemm_conv_ab.txt
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  launches %d" % (d["ms_per_step"], d["gpu_launches_per_step"]), {k[:24]: (round(v["ms_per_step"],1), round(v["tflops"],1)) for k,v in d.get("roofline",{}).get("kernels",{}).items()}, d.get("losses", d.get("loss")))'
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae.json 2> gpurun_out/${T}_vqvae.err; python -c "$P" gpurun_out/${T}_vqvae.json; tail -n 3 gpurun_out/${T}_vqvae.err | cut -c1-300
TTTS_GEMM_DEINT=0 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_nodeint.json 2> gpurun_out/${T}_vqvae_nodeint.err; python -c "$P" gpurun_out/${T}_vqvae_nodeint.json
