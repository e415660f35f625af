# Round 2, GPU call M: first hardware run of the diffusion mel-refiner step (BASELINE config 5): kernel / graph / step parity tests, first timing
# at batch 32 x 1024 frames with the REAL reference on the host cores beside it, launch list of the step at batch 8.
mkdir -p gpurun_out
T=r2m
timeout 900 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -rf > gpurun_out/${T}_pytest_diffusion.log 2>&1
echo "== diffusion rc=$?"; tail -4 gpurun_out/${T}_pytest_diffusion.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest_diffusion.log | cut -c1-300 | head -30
timeout 900 python tools/diffusion_step_bench.py 32 2 --cpu > gpurun_out/${T}_diffusion_step.json 2> gpurun_out/${T}_diffusion_step.err; cut -c1-2500 gpurun_out/${T}_diffusion_step.json; tail -3 gpurun_out/${T}_diffusion_step.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${T}_launches_diffusion_b8.csv python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_diffusion_ncu.log 2>&1; tail -1 gpurun_out/${T}_diffusion_ncu.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/${T}_launches_diffusion_b8.csv > gpurun_out/${T}_launches_diffusion_b8_summary.txt 2>&1; head -30 gpurun_out/${T}_launches_diffusion_b8_summary.txt
