mkdir -p gpurun_out
timeout 600 python tools/gemm_check.py > gpurun_out/r1i_gemm_check.txt 2>&1; grep -c PASS gpurun_out/r1i_gemm_check.txt; grep "FAIL\|TIME\|cuBLAS\|Error\|error" gpurun_out/r1i_gemm_check.txt | head -40
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gpt.py -m gpu -x -q > gpurun_out/r1i_pytest.log 2>&1; tail -5 gpurun_out/r1i_pytest.log
timeout 300 python tools/gemm_step_prof.py > gpurun_out/r1i_gemm_step_table.txt 2>&1; cat gpurun_out/r1i_gemm_step_table.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1i_bench.json 2>gpurun_out/r1i_bench.err; cut -c1-330 gpurun_out/r1i_bench.json; tail -3 gpurun_out/r1i_bench.err
