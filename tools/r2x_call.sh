# Round 2, GPU call X: launch list of the cfg2 step (BASELINE config 2: 12L / d512 / batch 8 / codes 512; 4.6 ms per step) -- which kernels are latency?
mkdir -p gpurun_out
T=r2x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --workload cfg2 --steps 3 --warmup 2 --profile-run --no-cpu-baseline --no-e2e --no-vq-encode --no-vqvae-step --no-diffusion-step > gpurun_out/${T}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_cfg2.csv > gpurun_out/${T}_launches_cfg2_summary.txt 2>&1; head -n 40 gpurun_out/${T}_launches_cfg2_summary.txt
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2x_launches_cfg2.csv')) if len(r)>5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, ... Metric Value last
by=collections.defaultdict(list)
for r in rows:
    by[(r[4][:50], r[8], r[7])].append(float(r[-1].replace(',','')))
out=sorted(by.items(), key=lambda kv:-sum(kv[1]))[:30]
for (k,g,b),v in out: print('%-52s grid %-18s block %-14s n %4d  avg %8.1f us  total %9.1f' % (k,g,b,len(v),sum(v)/len(v)/1e3 if max(v)>1e4 else sum(v)/len(v), sum(v)))
PY
rm -f gpurun_out/${T}_launches_cfg2.csv
