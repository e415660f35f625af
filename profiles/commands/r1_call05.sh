mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1k_2gpu_$tag.json 2> gpurun_out/r1k_2gpu_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r1k_2gpu_$tag.json') if l.startswith('{')][-1])
    print('$tag', 'ms/step %.2f'%d['ms_per_step'], 'frames/s %.0f'%d['value'], d['clocks'], 'gemm', '%.0f'%d['roofline']['achieved'])
except Exception as e:
    print('$tag', 'FAILED', e)
PY
}
run overlap6 X=1
run nooverlap TTTS_COMM_CHUNKS=0
run overlap6_cta8 NCCL_MAX_CTAS=8
run overlap6_cta2 NCCL_MAX_CTAS=2
run overlap2 TTTS_COMM_CHUNKS=2
timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1k_1gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r1k_1gpu.json
tail -2 gpurun_out/r1k_2gpu_overlap6.err
