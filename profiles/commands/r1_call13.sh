mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gpt.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ddp_check.py > gpurun_out/r1s_ddp_check_2gpu.log 2>&1; tail -4 gpurun_out/r1s_ddp_check_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1s_bench_2gpu.json 2> gpurun_out/r1s_bench_2gpu.err; cut -c1-420 gpurun_out/r1s_bench_2gpu.json; tail -2 gpurun_out/r1s_bench_2gpu.err
