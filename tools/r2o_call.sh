# Round 2, GPU call O: strided input gradients / transposed convolutions by phase decomposition on the forward kernels; opt-in tensor-core
# training convolutions (TTTS_TRAIN_TC=1); launch lists of both training steps after the conv backward rework.
mkdir -p gpurun_out
T=r2o
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -30
TTTS_TRAIN_TC=1 timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest_train_tc.log 2>&1
echo "== pytest TRAIN_TC rc=$?"; tail -4 gpurun_out/${T}_pytest_train_tc.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest_train_tc.log | cut -c1-300 | head -20
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  TFLOP/s %.1f  launches %d" % (d["ms_per_step"], d["step_tflops"], d["gpu_launches_per_step"]), {k: round(v["ms_per_step"],1) for k,v in d.get("roofline",{}).get("kernels",{}).items()}, d.get("losses", d.get("loss")))'
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_phase.json 2> gpurun_out/${T}_vqvae_phase.err; python -c "$P" gpurun_out/${T}_vqvae_phase.json
TTTS_TRAIN_TC=1 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_phase_tc.json 2> gpurun_out/${T}_vqvae_phase_tc.err; python -c "$P" gpurun_out/${T}_vqvae_phase_tc.json
timeout 600 python tools/diffusion_step_bench.py 32 2 > gpurun_out/${T}_diffusion.json 2> gpurun_out/${T}_diffusion.err; python -c "$P" gpurun_out/${T}_diffusion.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${T}_launches_diffusion_b32.csv python tools/diffusion_step_bench.py 32 1 > gpurun_out/${T}_diffusion_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_diffusion_b32.csv > gpurun_out/${T}_launches_diffusion_b32_summary.txt 2>&1; head -24 gpurun_out/${T}_launches_diffusion_b32_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/${T}_launches_vqvae_b16.csv python tools/vqvae_step_bench.py 16 1 > gpurun_out/${T}_vqvae_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_vqvae_b16.csv > gpurun_out/${T}_launches_vqvae_b16_summary.txt 2>&1; head -30 gpurun_out/${T}_launches_vqvae_b16_summary.txt
rm -f gpurun_out/${T}_launches_vqvae_b16.csv
