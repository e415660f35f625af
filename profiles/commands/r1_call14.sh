mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r1t_bench_8gpu.json 2> gpurun_out/r1t_bench_8gpu.err; cut -c1-420 gpurun_out/r1t_bench_8gpu.json; tail -3 gpurun_out/r1t_bench_8gpu.err
