mkdir -p gpurun_out
T=r2d
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/${T}_pytest_attn.log 2>&1; tail -2 gpurun_out/${T}_pytest_attn.log | cut -c1-300
timeout 200 python tools/attn_prof.py 2>&1 | tee gpurun_out/${T}_attn_ab.txt
TTTS_CONV_TC_FLAGS=0 timeout 200 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_tcs" > gpurun_out/${T}_pytest_tcs_baseoff.log 2>&1; tail -4 gpurun_out/${T}_pytest_tcs_baseoff.log | cut -c1-400
TTTS_CONV_TC_FLAGS=1 timeout 200 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_tcs" > gpurun_out/${T}_pytest_tcs_nobaseoff.log 2>&1; tail -4 gpurun_out/${T}_pytest_tcs_nobaseoff.log | cut -c1-400
grep -h "AssertionError\|assert " gpurun_out/${T}_pytest_tcs_*.log | head -12 | cut -c1-300
(TTTS_CONV_TC=0 timeout 200 python tools/enc_tc_check.py; TTTS_CONV_TC=1 TTTS_CONV_TC_FLAGS=0 timeout 200 python tools/enc_tc_check.py; TTTS_CONV_TC=1 TTTS_CONV_TC_FLAGS=1 timeout 200 python tools/enc_tc_check.py) 2>&1 | grep -v Warning | tee gpurun_out/${T}_enc_tc_check.txt | cut -c1-400
timeout 400 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_encoder.py -m gpu -q -x > gpurun_out/${T}_pytest_gpt_enc.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpt_enc.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vq-encode --no-vqvae-step > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-900 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
