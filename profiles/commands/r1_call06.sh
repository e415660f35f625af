mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_vq_mel.py -m gpu -x -q > gpurun_out/r1l_pytest_enc.log 2>&1; tail -3 gpurun_out/r1l_pytest_enc.log
for pipe in 1 0; do TTTS_CONV_PIPE=$pipe timeout 300 python tools/vq_encode_bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PIPE=$pipe', {k:d[k] for k in ('encode_ms_per_batch','encode_msamples_per_s','encode_graphed_ms_per_batch')})"; done
for pf in 0 1; do ONLY=proj_resid,fc2_resid,dgrad_pr_dgelu TTTS_GEMM_L2PF=$pf timeout 300 python tools/gemm_step_prof.py 2>&1 | sed "s/^/L2PF=$pf /"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1l_launches_vqenc.csv python tools/vq_encode_bench.py > /dev/null 2>&1; python tools/summarize_launches.py gpurun_out/r1l_launches_vqenc.csv | head -8
