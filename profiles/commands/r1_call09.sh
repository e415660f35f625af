mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vq_mel.py tests/test_gpu_encoder.py -m gpu -x -q > gpurun_out/r1o_pytest_vqmel.log 2>&1; tail -2 gpurun_out/r1o_pytest_vqmel.log
rm -f /tmp/kernels_ab.npz
(TTTS_STFT_V1=1 TTTS_VQ_V1=1 timeout 300 python tools/kernels_ab.py; timeout 300 python tools/kernels_ab.py) 2>&1 | tee gpurun_out/r1o_kernels_ab.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r1o_launches_step.csv python bench.py --steps 2 --warmup 1 --profile-run --no-cpu-baseline --no-vq-encode --no-e2e > gpurun_out/r1o_ncu_bench.log 2>&1; python tools/summarize_launches.py gpurun_out/r1o_launches_step.csv 2>/dev/null | head -20
