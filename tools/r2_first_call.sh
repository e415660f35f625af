# First GPU call of round 2: everything that was written at the end of round 1 without hardware (GPU budget spent), in order of risk.
# Usage:  /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
mkdir -p gpurun_out
# 1. regression: the default suite (must stay 69+ green), smoke
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2a_smoke.log 2>&1; tail -1 gpurun_out/r2a_smoke.log
# 2. KV-cache decode (gpt_decode.cu; validated on the CPU emulation only)
TTTS_KV_TEST=1 timeout 600 python -m pytest tests/test_gpu_gpt.py -m gpu -q -k "kv" > gpurun_out/r2a_pytest_kv.log 2>&1; tail -15 gpurun_out/r2a_pytest_kv.log
timeout 600 python tools/decode_bench.py 1 32 256 > gpurun_out/r2a_decode_b1.json 2> gpurun_out/r2a_decode_b1.err; cat gpurun_out/r2a_decode_b1.json; tail -2 gpurun_out/r2a_decode_b1.err
timeout 600 python tools/decode_bench.py 8 32 256 > gpurun_out/r2a_decode_b8.json 2> gpurun_out/r2a_decode_b8.err; cat gpurun_out/r2a_decode_b8.json; tail -2 gpurun_out/r2a_decode_b8.err
# 3. split-reduction convolution (conv1d_split.cu; CPU emulation only): kernel parity, whole-encoder parity with it on, A/B of the encode
TTTS_SPLIT_TEST=1 timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_split" > gpurun_out/r2a_pytest_split.log 2>&1; tail -8 gpurun_out/r2a_pytest_split.log
TTTS_CONV_SPLIT=1 timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/r2a_pytest_enc_split.log 2>&1; tail -8 gpurun_out/r2a_pytest_enc_split.log
(ONLY=enc timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_CONV_SPLIT=1 timeout 200 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/r2a_conv_split_ab.txt
TTTS_ENC_OVERLAP=1 timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/r2a_pytest_enc_overlap.log 2>&1; tail -4 gpurun_out/r2a_pytest_enc_overlap.log
(ONLY=enc TTTS_ENC_OVERLAP=1 timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_ENC_OVERLAP=1 TTTS_CONV_SPLIT=1 timeout 200 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/r2a_enc_overlap_ab.txt
TTTS_BWD_TEST=1 timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_backward or training_graph or generator_step or full_train_step" > gpurun_out/r2a_pytest_convbwd.log 2>&1; tail -6 gpurun_out/r2a_pytest_convbwd.log
timeout 600 python tools/vqvae_step_bench.py 8 > gpurun_out/r2a_vqvae_step_b8.json 2> gpurun_out/r2a_vqvae_step_b8.err; cat gpurun_out/r2a_vqvae_step_b8.json; tail -2 gpurun_out/r2a_vqvae_step_b8.err
# 4. split-bf16 tcgen05 convolution (conv1d_tc.cu; desk-checked + index emulation only).  Own timeout: a hang must not take the box.
TTTS_CONV_TC=1 timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_tc" > gpurun_out/r2a_pytest_convtc.log 2>&1; tail -15 gpurun_out/r2a_pytest_convtc.log
(ONLY=enc timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_CONV_TC=1 timeout 200 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/r2a_conv_tc_ab.txt
TTTS_CONV_TC=1 timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/r2a_pytest_enc_tc.log 2>&1; tail -8 gpurun_out/r2a_pytest_enc_tc.log
