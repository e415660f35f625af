# Round 2, GPU call AH: launch list of ONE routed convolution layer (tools/gemm_conv_prof.py): where the non-GEMM time of the route goes.
mkdir -p gpurun_out
T=r2ah
ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_conv.csv python tools/gemm_conv_prof.py > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2ah_launches_conv.csv')) if len(r)>5 and r[0].isdigit()]
out=open('gpurun_out/r2ah_launches_conv_list.txt','w')
for r in rows:
    line='%4s %-70s grid %-16s block %-12s %10.1f us' % (r[0], r[4][:70], r[8], r[7], float(r[-1].replace(',',''))/1e3)
    print(line); out.write(line+'\n')
PY
rm -f gpurun_out/${T}_launches_conv.csv
