# Round 2, GPU call U: decode step with the row-op and activation fused into the sweeps (5 launches per layer): parity, rates, launch list.
mkdir -p gpurun_out
T=r2u
timeout 600 python -m pytest tests/test_gpu_gpt.py -m gpu -q -rf -k "kv or inference" > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -n 3 gpurun_out/${T}_pytest.log | cut -c1-300
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -20
for B in 1 8; do timeout 300 python tools/decode_bench.py $B 32 256 > gpurun_out/${T}_decode_b$B.json 2> gpurun_out/${T}_decode_b$B.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/${T}_decode_b$B.json') if l.startswith('{')][-1]); print('B=$B uncached %.3f ms  eager %.3f  graph %.3f ms/code  frac_hbm %.3f' % (d['uncached_ms_per_code'], d['cached_eager_ms_per_code'], d['cached_graph_ms_per_code'], d['cached_graph_frac_hbm']))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dec_ -s 2000 -c 4000 --csv --log-file gpurun_out/${T}_launches_decode_b1.csv python tools/decode_bench.py 1 32 64 > gpurun_out/${T}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_decode_b1.csv > gpurun_out/${T}_launches_decode_b1_summary.txt 2>&1; head -n 10 gpurun_out/${T}_launches_decode_b1_summary.txt
rm -f gpurun_out/${T}_launches_decode_b1.csv
