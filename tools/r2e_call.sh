mkdir -p gpurun_out
T=r2e
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/${T}_pytest_attn.log 2>&1; tail -2 gpurun_out/${T}_pytest_attn.log | cut -c1-300
timeout 200 python tools/attn_prof.py 2>&1 | tee gpurun_out/${T}_attn_ab.txt
timeout 200 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "conv1d_tcs" > gpurun_out/${T}_pytest_tcs.log 2>&1; tail -4 gpurun_out/${T}_pytest_tcs.log | cut -c1-400
(TTTS_CONV_TC=0 timeout 200 python tools/enc_tc_check.py; TTTS_CONV_TC=1 timeout 200 python tools/enc_tc_check.py) 2>&1 | grep -v Warning | tee gpurun_out/${T}_enc_tc_check.txt | cut -c1-400
TTTS_CONV_TC=1 ONLY=enc ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_vqenc_tc.csv python tools/kernels_ab.py > gpurun_out/${T}_ncu_enc.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_vqenc_tc.csv > gpurun_out/${T}_launches_vqenc_tc_summary.txt 2>&1; head -24 gpurun_out/${T}_launches_vqenc_tc_summary.txt
for f in gpt encoder vq_mel; do
  timeout 500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rf > gpurun_out/${T}_pytest_$f.log 2>&1; echo "== $f rc=$?"; tail -3 gpurun_out/${T}_pytest_$f.log | cut -c1-300
done
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vq-encode --no-vqvae-step > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-900 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
