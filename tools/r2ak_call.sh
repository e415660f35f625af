# Round 2, GPU call AK: validation of the committed state (GPU suite, smoke, default bench line) + launch list of the VQ-VAE-GAN step at B = 64.
mkdir -p gpurun_out
T=r2ak
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "== pytest -m gpu rc=$? ($(( $(date +%s) - S )) s)"; tail -n 1 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
grep -h "^FAILED\|^ERROR" gpurun_out/${T}_pytest_gpu.log | cut -c1-300 | head -20
S=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "== smoke rc=$? ($(( $(date +%s) - S )) s)"; tail -n 1 gpurun_out/${T}_smoke.log | cut -c1-400
S=$(date +%s); timeout 1500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "== bench rc=$? ($(( $(date +%s) - S )) s)"; tail -n 2 gpurun_out/${T}_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ak_bench.json') if l.startswith('{')][-1])
print('value %.0f %s  ms/step %.2f  e2e %.0f  roofline frac %.3f  clocks %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('clocks')))
for k in ('cfg2','vq_encode','vqvae_step','diffusion_step'):
    v=d.get(k,{}); print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','frames_per_s','msamples_per_s','error','encode_ms_per_batch','encode_graphed_ms_per_batch')})
print('vqvae families', {k[:20]: (round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in d['vqvae_step'].get('roofline',{}).get('kernels',{}).items()})
PY
S=$(date +%s)
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 17500 --csv --log-file gpurun_out/${T}_launches_vqvae_b64.csv python tools/vqvae_step_bench.py 64 1 > gpurun_out/${T}_vqvae_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_vqvae_b64.csv > gpurun_out/${T}_launches_vqvae_b64_summary.txt 2>&1; head -n 36 gpurun_out/${T}_launches_vqvae_b64_summary.txt | cut -c1-120
rm -f gpurun_out/${T}_launches_vqvae_b64.csv
echo "== launch list ($(( $(date +%s) - S )) s)"
