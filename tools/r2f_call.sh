mkdir -p gpurun_out
T=r2f
timeout 200 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_tcs" > gpurun_out/${T}_pytest_tcs.log 2>&1; tail -3 gpurun_out/${T}_pytest_tcs.log | cut -c1-400
(TTTS_CONV_TC=1 timeout 200 python tools/enc_tc_check.py) 2>&1 | grep -v Warning | tee gpurun_out/${T}_enc_tc_check.txt | cut -c1-400
TTTS_CONV_TC=1 ONLY=enc ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_vqenc_tc.csv python tools/kernels_ab.py > gpurun_out/${T}_ncu_enc.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_vqenc_tc.csv > gpurun_out/${T}_launches_vqenc_tc_summary.txt 2>&1; head -12 gpurun_out/${T}_launches_vqenc_tc_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 4 -c 1 -o gpurun_out/${T}_gemm_qkv -f python tools/gemm_prof.py > gpurun_out/${T}_ncu_gemm.log 2>&1; tail -2 gpurun_out/${T}_ncu_gemm.log
TTTS_CONV_TC=1 ONLY=enc ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv1d_tcs -s 200 -c 3 -o gpurun_out/${T}_conv_tcs -f python tools/kernels_ab.py > gpurun_out/${T}_ncu_tcs.log 2>&1; tail -2 gpurun_out/${T}_ncu_tcs.log
