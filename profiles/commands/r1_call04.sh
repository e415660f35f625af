mkdir -p gpurun_out
timeout 600 python tools/gemm_check.py > gpurun_out/r1j_gemm_check.txt 2>&1; grep -c PASS gpurun_out/r1j_gemm_check.txt; grep "FAIL\|Error\|error" gpurun_out/r1j_gemm_check.txt | head
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1j_pytest.log 2>&1; tail -4 gpurun_out/r1j_pytest.log
timeout 300 python tools/gemm_step_prof.py > gpurun_out/r1j_gemm_step_table.txt 2>&1; cat gpurun_out/r1j_gemm_step_table.txt
ONLY=proj_resid,fc2_resid,dgrad_pr_dgelu TTTS_GEMM_L2PF=0 timeout 300 python tools/gemm_step_prof.py 2>&1 | sed 's/^/L2PF=0 /'
timeout 300 python tools/attn_prof.py 2>&1 | tail -3
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1j_bench.json 2>gpurun_out/r1j_bench.err; cut -c1-330 gpurun_out/r1j_bench.json; tail -3 gpurun_out/r1j_bench.err
