# Round 2, GPU call C: attention math rewrite (15-bit hash fields, packed masks), kink-aware encoder tests, bench with vqvae_step + real reference arm,
# launch list of the VQ-VAE-GAN step, ncu --set full of the new attention kernels, compute-sanitizer.
mkdir -p gpurun_out
T=r2c
for f in kernels gpt encoder; do
  timeout 700 python -m pytest tests/test_gpu_$f.py -m gpu -q -rf -x > gpurun_out/${T}_pytest_$f.log 2>&1
  echo "== $f rc=$?"; tail -3 gpurun_out/${T}_pytest_$f.log | cut -c1-300
done
grep -h "^FAILED\|^ERROR" gpurun_out/${T}_pytest_*.log | cut -c1-400
timeout 200 python tools/attn_prof.py 2>&1 | tee gpurun_out/${T}_attn_ab.txt
ONLY_P=0.1 ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc4 -c 1 -o gpurun_out/${T}_attn_fwd -f python tools/attn_prof.py > gpurun_out/${T}_ncu_fwd.log 2>&1; tail -2 gpurun_out/${T}_ncu_fwd.log
ONLY_P=0.1 ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc4 -c 1 -o gpurun_out/${T}_attn_bwd -f python tools/attn_prof.py > gpurun_out/${T}_ncu_bwd.log 2>&1; tail -2 gpurun_out/${T}_ncu_bwd.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-1500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; cut -c1-2500 gpurun_out/${T}_ref.json; tail -3 gpurun_out/${T}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${T}_launches_vqvae_b8.csv python tools/vqvae_step_bench.py 8 1 > gpurun_out/${T}_vqvae_ncu.log 2>&1; tail -1 gpurun_out/${T}_vqvae_ncu.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/${T}_launches_vqvae_b8.csv > gpurun_out/${T}_launches_vqvae_b8_summary.txt 2>&1; head -25 gpurun_out/${T}_launches_vqvae_b8_summary.txt
for tool in racecheck synccheck memcheck; do
  timeout 280 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_epilogues or attention_fwd_bwd or layernorm or cross_entropy or adamw" > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "== sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/${T}_sanitizer_$tool.log | tail -3
done
