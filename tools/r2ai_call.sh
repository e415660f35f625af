# Round 2, GPU call AI: layout kernels on 64 x 64 tiles with packed stores, leaky-ReLU derivative folded into cl_unpack, one split of dy for both
# gradients of a "same" convolution, 16-byte-load bias gradient, leaner weight preparation: parity + A/B.
mkdir -p gpurun_out
T=r2ai
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -1 gpurun_out/${T}_pytest.log | cut -c1-300; grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-250 | head -20
timeout 200 python tools/gemm_conv_prof.py | tee gpurun_out/${T}_gemm_conv_ab.txt
timeout 200 python tools/gemm_conv_prof.py 32 512 1024 512 3 1 1 1 | tee -a gpurun_out/${T}_gemm_conv_ab.txt
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  launches %d" % (d["ms_per_step"], d["gpu_launches_per_step"]), {k[:24]: (round(v["ms_per_step"],1), round(v["tflops"],1)) for k,v in d.get("roofline",{}).get("kernels",{}).items()}, d.get("losses", d.get("loss")))'
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae.json 2> gpurun_out/${T}_vqvae.err; python -c "$P" gpurun_out/${T}_vqvae.json; tail -n 3 gpurun_out/${T}_vqvae.err | cut -c1-300
TTTS_GEMM_NARROW=1 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_narrow.json 2> gpurun_out/${T}_vqvae_narrow.err; python -c "$P" gpurun_out/${T}_vqvae_narrow.json
timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion.json 2> gpurun_out/${T}_diffusion.err; python -c "$P" gpurun_out/${T}_diffusion.json; tail -n 3 gpurun_out/${T}_diffusion.err | cut -c1-300
ITERS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_conv.csv python tools/gemm_conv_prof.py > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2ai_launches_conv.csv')) if len(r)>5 and r[0].isdigit()]
out=open('gpurun_out/r2ai_launches_conv_list.txt','w')
for r in rows:
    line='%4s %-70s grid %-16s block %-12s %10.1f us' % (r[0], r[4][:70], r[8], r[7], float(r[-1].replace(',',''))/1e3)
    out.write(line+'\n')
    if 'ttts' in r[4]: print(line)
PY
rm -f gpurun_out/${T}_launches_conv.csv
