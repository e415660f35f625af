mkdir -p gpurun_out
T=r2h
timeout 90 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x -k "conv1d_tcs" > gpurun_out/${T}_pytest_tcs.log 2>&1; rc=$?; tail -3 gpurun_out/${T}_pytest_tcs.log | cut -c1-400
if [ $rc -ne 0 ]; then echo "conv1d_tcs unit tests failed (rc=$rc): stopping"; grep -h "Error\|assert" gpurun_out/${T}_pytest_tcs.log | head; exit 1; fi
(TTTS_CONV_TC=1 timeout 120 python tools/enc_tc_check.py) 2>&1 | grep -v Warning | tee gpurun_out/${T}_enc_tc_check.txt | cut -c1-400
TTTS_CONV_TC=1 ONLY=enc ITERS=2 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_vqenc_tc.csv python tools/kernels_ab.py > gpurun_out/${T}_ncu_enc.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_vqenc_tc.csv > gpurun_out/${T}_launches_vqenc_tc_summary.txt 2>&1; head -8 gpurun_out/${T}_launches_vqenc_tc_summary.txt
TTTS_CONV_TC=1 timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest_encoder_tc1.log 2>&1; echo "== encoder (TTTS_CONV_TC=1) rc=$?"; tail -3 gpurun_out/${T}_pytest_encoder_tc1.log | cut -c1-300
TTTS_CONV_TC=0 timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest_encoder_tc0.log 2>&1; echo "== encoder (TTTS_CONV_TC=0) rc=$?"; tail -3 gpurun_out/${T}_pytest_encoder_tc0.log | cut -c1-300
TTTS_CONV_TC=1 ONLY=enc ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tcs -s 60 -c 2 -o gpurun_out/${T}_conv_tcs -f python tools/kernels_ab.py > gpurun_out/${T}_ncu_tcs.log 2>&1; tail -2 gpurun_out/${T}_ncu_tcs.log
