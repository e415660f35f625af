mkdir -p gpurun_out
T=r2k
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-vq-encode --no-vqvae-step"
for c in 6 12 24; do
  TTTS_COMM_CHUNKS=$c timeout 300 $TR --master-port 2952$((c % 10)) $B > gpurun_out/${T}_bench_2gpu_chunks$c.json 2> gpurun_out/${T}_bench_2gpu_chunks$c.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_2gpu_chunks$c.json') if l.startswith('{')][-1])
print('chunks=$c N=2 ms/step %.2f  gemm ms %.2f  p10 %.2f median %.2f'%(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['step_ms_rank0']['p10'], d['step_ms_rank0']['median']))"
done
timeout 300 python bench.py --gpus 1 --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-vq-encode --no-vqvae-step > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_1gpu.json') if l.startswith('{')][-1])
print('1 GPU ms/step %.2f gemm %.2f p10 %.2f median %.2f'%(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['step_ms_rank0']['p10'], d['step_ms_rank0']['median']))"
