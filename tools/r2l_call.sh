mkdir -p gpurun_out
T=r2l
timeout 120 python -m pytest tests/test_gpu_vq_mel.py -m gpu -q -x -k "tensor_core" > gpurun_out/${T}_pytest_vqtc.log 2>&1; rc=$?; tail -3 gpurun_out/${T}_pytest_vqtc.log | cut -c1-400
if [ $rc -ne 0 ]; then echo "vq tc tests failed (rc=$rc)"; grep -h "Error\|assert\|error" gpurun_out/${T}_pytest_vqtc.log | head; exit 1; fi
(ONLY=vq TTTS_VQ_TC=0 timeout 120 python tools/kernels_ab.py; ONLY=vq TTTS_VQ_TC=1 timeout 120 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/${T}_vq_ab.txt | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_vq_mel.py tests/test_gpu_encoder.py -m gpu -q > gpurun_out/${T}_pytest_vq_enc.log 2>&1; tail -2 gpurun_out/${T}_pytest_vq_enc.log | cut -c1-300
ONLY=vq ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_tc_ -c 2 -s 6 -o gpurun_out/${T}_vq_tc -f python tools/kernels_ab.py > gpurun_out/${T}_ncu_vq.log 2>&1; tail -2 gpurun_out/${T}_ncu_vq.log
