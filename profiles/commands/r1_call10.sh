mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/vq_encode_bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('encode_ms_per_batch','encode_msamples_per_s','encode_graphed_ms_per_batch')})"
ONCE=1 ONLY=proj_resid,wgrad_proj,fc1_gelu timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm2 -f -o gpurun_out/r1p_gemm_epi python tools/gemm_step_prof.py > gpurun_out/r1p_ncu_gemm.log 2>&1; tail -3 gpurun_out/r1p_ncu_gemm.log
timeout 300 python tools/gemm_step_prof.py 2>&1 | tee gpurun_out/r1p_gemm_step_table.txt
