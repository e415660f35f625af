# Round 2, GPU call AG (2 GPUs): the data-parallel step on the committed state, launched the way the driver launches it.
mkdir -p gpurun_out
T=r2ag
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
echo "== rc=$?"; tail -n 3 gpurun_out/${T}_bench_2gpu.err | cut -c1-300
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2ag_bench_2gpu.json') if l.startswith('{')][-1]); print('n_gpus', d['n_gpus'], 'value %.0f' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], d['clocks'], d['config']['parallelism'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/${T}_ref_2gpu.json 2> gpurun_out/${T}_ref_2gpu.err
echo "== reference arm under torchrun rc=$?"; cut -c1-300 gpurun_out/${T}_ref_2gpu.json
