# Round 2, GPU call V: end-of-round validation -- the whole GPU suite, smoke(), the default bench line (both arms), decode rates.
mkdir -p gpurun_out
T=r2v
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "== pytest -m gpu rc=$? ($(( $(date +%s) - S )) s)"; tail -n 3 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
grep -h "^FAILED\|^ERROR" gpurun_out/${T}_pytest_gpu.log | cut -c1-300 | head -20
S=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "== smoke rc=$? ($(( $(date +%s) - S )) s)"; tail -n 1 gpurun_out/${T}_smoke.log | cut -c1-400
S=$(date +%s); timeout 1500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "== bench rc=$? ($(( $(date +%s) - S )) s)"; cut -c1-700 gpurun_out/${T}_bench.json; tail -n 3 gpurun_out/${T}_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v_bench.json') if l.startswith('{')][-1])
print('value %.0f %s  ms/step %.2f  e2e %.0f  roofline frac %.3f  clocks %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('clocks')))
for k in ('cfg2','vq_encode','vqvae_step','diffusion_step'):
    v=d.get(k,{}); print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','frames_per_s','msamples_per_s','error','encode_ms','graph_ms','encode_graph_ms','msamples_per_s_graph')})
print('cpu_baseline', d.get('cpu_baseline'))
PY
S=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "== reference arm rc=$? ($(( $(date +%s) - S )) s)"; cut -c1-600 gpurun_out/${T}_ref.json
for B in 1 8; do timeout 300 python tools/decode_bench.py $B 32 256 > gpurun_out/${T}_decode_b$B.json 2> gpurun_out/${T}_decode_b$B.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/${T}_decode_b$B.json') if l.startswith('{')][-1]); print('B=$B uncached %.3f ms  eager %.3f  graph %.3f ms/code  frac_hbm %.3f' % (d['uncached_ms_per_code'], d['cached_eager_ms_per_code'], d['cached_graph_ms_per_code'], d['cached_graph_frac_hbm']))"; done
