# Round 2, GPU call AC: attention tail kernels with batched loads (r2ab: the first version cost more than the tile it replaced): parity,
# per-kernel times (ncu launch list of tools/attn_prof.py), kernel and step A/B.
mkdir -p gpurun_out
T=r2ac
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rf -k "attention" > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -1 gpurun_out/${T}_pytest.log | cut -c1-300
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-250 | head -20
for v in 1 0; do
  echo "-- TTTS_ATTN_TAIL=$v"
  TTTS_ATTN_TAIL=$v ITERS=10 timeout 200 python tools/attn_prof.py 2>&1 | grep -v digest | tee -a gpurun_out/${T}_attn_ab.txt
  TTTS_ATTN_TAIL=$v ITERS=10 timeout 200 python tools/attn_prof.py 8 644 8 2>&1 | grep -v digest | tee -a gpurun_out/${T}_attn_ab.txt
done
ONLY_P=0.1 ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_attn.csv python tools/attn_prof.py > gpurun_out/${T}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_attn.csv > gpurun_out/${T}_launches_attn_summary.txt 2>&1; head -16 gpurun_out/${T}_launches_attn_summary.txt | cut -c1-150
Bq="bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vq-encode --no-vqvae-step --no-diffusion-step"
for v in 1 0; do
  TTTS_ATTN_TAIL=$v timeout 400 python $Bq > gpurun_out/${T}_bench_tail$v.json 2> gpurun_out/${T}_bench_tail$v.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/${T}_bench_tail$v.json') if l.startswith('{')][-1]); print('tail=$v', 'ms/step %.2f' % d['ms_per_step'], 'p10 %.2f' % d['step_ms_rank0']['p10'], 'e2e %.2f' % d['e2e']['ms_per_step'], 'cfg2 %.3f' % d.get('cfg2',{}).get('ms_per_step',-1), d['clocks'])"
done
