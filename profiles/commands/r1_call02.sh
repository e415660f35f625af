mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gpt.py -m gpu -x -q > gpurun_out/r1h_pytest.log 2>&1; tail -3 gpurun_out/r1h_pytest.log
python tools/gemm_step_prof.py > gpurun_out/r1h_gemm_step_table.txt 2>&1; cat gpurun_out/r1h_gemm_step_table.txt
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1h_bench_lnfix.json 2>gpurun_out/r1h_bench.err; cut -c1-330 gpurun_out/r1h_bench_lnfix.json
ONCE=1 ONLY=fc1_gelu,proj_resid,dgrad_pr_dgelu,qkv_bf16 timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm2 -f -o gpurun_out/r1h_gemm_epi python tools/gemm_step_prof.py > gpurun_out/r1h_ncu_gemm.log 2>&1
ITERS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_fwd_tc4 --launch-skip 3 -c 1 -f -o gpurun_out/r1h_attn_fwd4 python tools/attn_prof.py > gpurun_out/r1h_ncu_attn_fwd.log 2>&1
ITERS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_tc4 --launch-skip 3 -c 1 -f -o gpurun_out/r1h_attn_bwd4 python tools/attn_prof.py > gpurun_out/r1h_ncu_attn_bwd.log 2>&1
ls -la gpurun_out
