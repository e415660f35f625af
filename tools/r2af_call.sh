# Round 2, GPU call AF: validation of the committed state -- the whole GPU suite, smoke(), the default bench line (both arms), the launch list of
# the headline step (cfg3) for round 2, per-layer A/B and `ncu --set full` of the tap-concatenated convolution GEMM.
mkdir -p gpurun_out
T=r2af
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "== pytest -m gpu rc=$? ($(( $(date +%s) - S )) s)"; tail -n 1 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
grep -h "^FAILED\|^ERROR" gpurun_out/${T}_pytest_gpu.log | cut -c1-300 | head -20
S=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "== smoke rc=$? ($(( $(date +%s) - S )) s)"; tail -n 1 gpurun_out/${T}_smoke.log | cut -c1-400
S=$(date +%s); timeout 1500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "== bench rc=$? ($(( $(date +%s) - S )) s)"; tail -n 3 gpurun_out/${T}_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2af_bench.json') if l.startswith('{')][-1])
print('value %.0f %s  ms/step %.2f  e2e %.0f  roofline frac %.3f  clocks %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('clocks')))
for k in ('cfg2','vq_encode','vqvae_step','diffusion_step'):
    v=d.get(k,{}); print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','frames_per_s','msamples_per_s','error','encode_ms_per_batch','encode_graphed_ms_per_batch')})
print('cpu_baseline', d.get('cpu_baseline'))
PY
S=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "== reference arm rc=$? ($(( $(date +%s) - S )) s)"; cut -c1-500 gpurun_out/${T}_ref.json
S=$(date +%s)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_step.csv python bench.py --steps 2 --warmup 1 --profile-run --no-cpu-baseline --no-e2e --no-vq-encode --no-vqvae-step --no-diffusion-step > gpurun_out/${T}_ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_step.csv > gpurun_out/${T}_launches_summary.txt 2>&1; head -n 24 gpurun_out/${T}_launches_summary.txt | cut -c1-130
rm -f gpurun_out/${T}_launches_step.csv
echo "== launch list ($(( $(date +%s) - S )) s)"
timeout 200 python tools/gemm_conv_prof.py | tee gpurun_out/${T}_gemm_conv_ab.txt
TTTS_GEMM_TAPCAT=0 timeout 200 python tools/gemm_conv_prof.py | tee -a gpurun_out/${T}_gemm_conv_ab.txt
TTTS_TRAIN_GEMM=0 timeout 200 python tools/gemm_conv_prof.py | tee -a gpurun_out/${T}_gemm_conv_ab.txt
ITERS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm -s 4 -c 6 -o gpurun_out/${T}_gemm_conv -f python tools/gemm_conv_prof.py > gpurun_out/${T}_ncu_gemm.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_gemm_conv.ncu-rep > gpurun_out/${T}_gemm_conv_ncu_full.txt 2>&1; head -n 12 gpurun_out/${T}_gemm_conv_ncu_full.txt | cut -c1-200
