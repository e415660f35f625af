mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1x_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1x_smoke.log 2>&1; tail -1 gpurun_out/r1x_smoke.log
timeout 900 python bench.py > gpurun_out/r1x_bench.json 2> gpurun_out/r1x_bench.err; cut -c1-330 gpurun_out/r1x_bench.json; tail -2 gpurun_out/r1x_bench.err
