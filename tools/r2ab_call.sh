# Round 2, GPU call AB: ragged attention tail (T mod 128 <= 16 query rows on CUDA cores, tile kernels stop at Tm): parity, kernel A/B, step A/B;
# the generalised GEMM-route convolution test at its own (small) shapes.
mkdir -p gpurun_out
T=r2ab
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gpt.py tests/test_gpu_diffusion.py -m gpu -q -rf -k "attention or gpt or tensor_core" > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-250 | head -20
for v in 1 0; do
  echo "-- TTTS_ATTN_TAIL=$v"
  TTTS_ATTN_TAIL=$v ITERS=10 timeout 200 python tools/attn_prof.py 2>&1 | grep -v digest | tee -a gpurun_out/${T}_attn_ab.txt
  TTTS_ATTN_TAIL=$v ITERS=10 timeout 200 python tools/attn_prof.py 8 644 8 2>&1 | grep -v digest | tee -a gpurun_out/${T}_attn_ab.txt
done
Bq="bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vq-encode --no-vqvae-step --no-diffusion-step"
for v in 1 0; do
  TTTS_ATTN_TAIL=$v timeout 400 python $Bq > gpurun_out/${T}_bench_tail$v.json 2> gpurun_out/${T}_bench_tail$v.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/${T}_bench_tail$v.json') if l.startswith('{')][-1]); print('tail=$v', 'ms/step %.2f' % d['ms_per_step'], 'p10 %.2f' % d['step_ms_rank0']['p10'], 'e2e %.2f' % d['e2e']['ms_per_step'], 'cfg2 %.3f' % d.get('cfg2',{}).get('ms_per_step',-1), d['clocks'])"
done
