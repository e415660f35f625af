mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
(ONLY=enc TTTS_CONV_DIRECT=0 timeout 200 python tools/kernels_ab.py; ONLY=enc timeout 200 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/r1v_conv_direct_ab.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1v_launches_vqenc.csv python tools/vq_encode_bench.py > /dev/null 2>&1; python tools/summarize_launches.py gpurun_out/r1v_launches_vqenc.csv 2>/dev/null | head -12
