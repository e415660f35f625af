mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1q_pytest.log 2>&1; tail -3 gpurun_out/r1q_pytest.log
for v in 4 5; do TTTS_ATTN_VER=$v ONLY_P=0.1 timeout 120 python tools/attn_prof.py 2>&1 | sed "s/^/VER=$v /"; done | tee gpurun_out/r1q_attn_ab.txt
rm -f /tmp/kernels_ab.npz
(ONLY=stft TTTS_STFT_V1=1 timeout 300 python tools/kernels_ab.py; ONLY=stft timeout 300 python tools/kernels_ab.py; ONLY=mel24 timeout 300 python tools/kernels_ab.py) 2>&1 | tee gpurun_out/r1q_kernels_ab.txt
for pdl in 0 1; do TTTS_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1q_bench_pdl$pdl.json 2> gpurun_out/r1q_bench_pdl$pdl.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r1q_bench_pdl$pdl.json') if l.startswith('{')][-1])
    print('PDL=$pdl', 'ms/step %.2f'%d['ms_per_step'], d['step_ms_rank0'], 'e2e %.2f'%d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], 'gemm %.0f'%d['roofline']['achieved'], 'cfg2 %.2f'%d['cfg2']['ms_per_step'])
except Exception as e:
    print('PDL=$pdl FAILED', e)
PY
done
