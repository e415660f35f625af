mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1r_pytest_gpu.log 2>&1; tail -4 gpurun_out/r1r_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1r_smoke.log 2>&1; tail -2 gpurun_out/r1r_smoke.log
timeout 900 python bench.py > gpurun_out/r1r_bench.json 2> gpurun_out/r1r_bench.err; cut -c1-600 gpurun_out/r1r_bench.json; tail -2 gpurun_out/r1r_bench.err
