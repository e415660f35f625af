# Round 2, GPU call AA: the split-bf16 GEMM route of the wide training convolutions generalised (stride, dilation, padding, leaky ReLU) and taken
# by the VQ-VAE-GAN tape too; parity + A/B + launch list of the VQ-VAE-GAN step.
mkdir -p gpurun_out
T=r2aa
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -30
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  TFLOP/s %.1f  launches %d" % (d["ms_per_step"], d["step_tflops"], d["gpu_launches_per_step"]), d.get("losses", d.get("loss")))'
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_gemm.json 2> gpurun_out/${T}_vqvae_gemm.err; python -c "$P" gpurun_out/${T}_vqvae_gemm.json; tail -n 3 gpurun_out/${T}_vqvae_gemm.err | cut -c1-300
TTTS_TRAIN_GEMM=0 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_fp32.json 2> gpurun_out/${T}_vqvae_fp32.err; python -c "$P" gpurun_out/${T}_vqvae_fp32.json
timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion.json 2> gpurun_out/${T}_diffusion.err; python -c "$P" gpurun_out/${T}_diffusion.json; tail -n 3 gpurun_out/${T}_diffusion.err | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/${T}_launches_vqvae_b16.csv python tools/vqvae_step_bench.py 16 1 > gpurun_out/${T}_vqvae_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_vqvae_b16.csv > gpurun_out/${T}_launches_vqvae_b16_summary.txt 2>&1; head -30 gpurun_out/${T}_launches_vqvae_b16_summary.txt
rm -f gpurun_out/${T}_launches_vqvae_b16.csv
