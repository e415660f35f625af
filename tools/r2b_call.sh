mkdir -p gpurun_out
T=r2b
timeout 400 python tools/bwd_debug.py > gpurun_out/${T}_bwd_debug.txt 2>&1; tail -70 gpurun_out/${T}_bwd_debug.txt | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest_encoder.log 2>&1; tail -8 gpurun_out/${T}_pytest_encoder.log | cut -c1-300
(ONLY=enc timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_CONV_SPLIT=1 timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_CONV_TC=1 timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_ENC_OVERLAP=1 timeout 200 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/${T}_conv_ab.txt | cut -c1-400
TTTS_CONV_TC=1 timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "golden or batch64" > gpurun_out/${T}_pytest_enc_tc.log 2>&1; tail -6 gpurun_out/${T}_pytest_enc_tc.log | cut -c1-300
TTTS_CONV_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "encoder_vs or batch64" > gpurun_out/${T}_pytest_enc_split.log 2>&1; tail -4 gpurun_out/${T}_pytest_enc_split.log | cut -c1-300
