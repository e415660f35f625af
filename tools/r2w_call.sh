# Round 2, GPU call W: split exact-fp32 VQ sweep for small N (the encode metric's N = 1 152): bit identity, timing, encode.
mkdir -p gpurun_out
T=r2w
timeout 600 python -m pytest tests/test_gpu_vq_mel.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -n 3 gpurun_out/${T}_pytest.log | cut -c1-300
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -20
(ONLY=vq TTTS_VQ_SPLIT=0 timeout 120 python tools/kernels_ab.py; ONLY=vq TTTS_VQ_SPLIT=1 timeout 120 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | grep "N=1152" | tee gpurun_out/${T}_vq_ab.txt | cut -c1-300
timeout 300 python tools/enc_tc_check.py > gpurun_out/${T}_enc_check.txt 2>&1; tail -n 3 gpurun_out/${T}_enc_check.txt | cut -c1-300
