# Round 2, GPU call A: everything written without hardware at the end of round 1, now un-gated, plus the new parity tests.
# Usage: gpurun --timeout 1500 -- 'bash tools/r2a_call.sh'
mkdir -p gpurun_out
T=r2a
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${T}_gpu.txt 2>&1
for f in kernels vq_mel gpt encoder; do
  timeout 700 python -m pytest tests/test_gpu_$f.py -m gpu -q -rf --durations=8 > gpurun_out/${T}_pytest_$f.log 2>&1
  echo "== $f rc=$?"; tail -4 gpurun_out/${T}_pytest_$f.log | cut -c1-300
done
grep -h "^FAILED\|^ERROR" gpurun_out/${T}_pytest_*.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 300 python tools/decode_bench.py 1 32 256 > gpurun_out/${T}_decode_b1.json 2> gpurun_out/${T}_decode_b1.err; cat gpurun_out/${T}_decode_b1.json; tail -2 gpurun_out/${T}_decode_b1.err
timeout 300 python tools/decode_bench.py 8 32 256 > gpurun_out/${T}_decode_b8.json 2> gpurun_out/${T}_decode_b8.err; cat gpurun_out/${T}_decode_b8.json; tail -2 gpurun_out/${T}_decode_b8.err
timeout 400 python tools/vqvae_step_bench.py 8 > gpurun_out/${T}_vqvae_step_b8.json 2> gpurun_out/${T}_vqvae_step_b8.err; cat gpurun_out/${T}_vqvae_step_b8.json; tail -2 gpurun_out/${T}_vqvae_step_b8.err
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_step_b64.json 2> gpurun_out/${T}_vqvae_step_b64.err; cat gpurun_out/${T}_vqvae_step_b64.json; tail -2 gpurun_out/${T}_vqvae_step_b64.err
(ONLY=enc timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_CONV_SPLIT=1 timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_CONV_TC=1 timeout 200 python tools/kernels_ab.py; ONLY=enc TTTS_ENC_OVERLAP=1 timeout 200 python tools/kernels_ab.py) 2>&1 | grep -v "^$" | tee gpurun_out/${T}_conv_ab.txt
TTTS_CONV_TC=1 timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "golden or batch64" > gpurun_out/${T}_pytest_enc_tc.log 2>&1; tail -6 gpurun_out/${T}_pytest_enc_tc.log | cut -c1-300
TTTS_CONV_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "encoder_vs or batch64" > gpurun_out/${T}_pytest_enc_split.log 2>&1; tail -4 gpurun_out/${T}_pytest_enc_split.log | cut -c1-300
