# Round 2, GPU call Z: cfg2 A/B: first-generation GEMM kernel (128 x 128/256 tiles, single CTA, no epilogue overlap) on the latency-bound step.
mkdir -p gpurun_out
T=r2z
B="bench.py --workload cfg2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-vq-encode --no-vqvae-step --no-diffusion-step"
for v in default legacy; do
  if [ $v = legacy ]; then export TTTS_GEMM_LEGACY=1; fi
  timeout 300 python $B > gpurun_out/${T}_cfg2_$v.json 2> gpurun_out/${T}_cfg2_$v.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/${T}_cfg2_$v.json') if l.startswith('{')][-1]); print('$v', 'ms/step %.3f' % d['ms_per_step'], 'gemm ms %.3f' % d['roofline']['kernel_ms_per_step'], 'p10 %.3f' % d['step_ms_rank0']['p10'])"
done
