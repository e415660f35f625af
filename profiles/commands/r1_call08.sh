mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
ONLY=stft ITERS=1 timeout 300 $N -k regex:stft_mel_r16 -s 2 -c 1 -o gpurun_out/r1n_stft_r16 python tools/kernels_ab.py > gpurun_out/r1n_ncu_stft.log 2>&1; tail -2 gpurun_out/r1n_ncu_stft.log
ONLY=vq ITERS=1 timeout 300 $N -k regex:vq_argmin_pipe -s 24 -c 1 -o gpurun_out/r1n_vq_pipe python tools/kernels_ab.py > gpurun_out/r1n_ncu_vq.log 2>&1; tail -2 gpurun_out/r1n_ncu_vq.log
ONLY_P=0.1 ITERS=1 timeout 300 $N -k regex:attn_fwd_tc4 -s 2 -c 1 -o gpurun_out/r1n_attn_fwd5 python tools/attn_prof.py > gpurun_out/r1n_ncu_attn_fwd.log 2>&1; tail -2 gpurun_out/r1n_ncu_attn_fwd.log
ONLY_P=0.1 ITERS=1 timeout 300 $N -k regex:attn_bwd_tc4 -s 2 -c 1 -o gpurun_out/r1n_attn_bwd5 python tools/attn_prof.py > gpurun_out/r1n_ncu_attn_bwd.log 2>&1; tail -2 gpurun_out/r1n_ncu_attn_bwd.log
timeout 400 $N -k regex:ln_bwd_kernel -s 30 -c 1 -o gpurun_out/r1n_ln_bwd python bench.py --steps 1 --warmup 1 --profile-run --no-cpu-baseline --no-vq-encode --no-e2e > gpurun_out/r1n_ncu_ln.log 2>&1; tail -2 gpurun_out/r1n_ncu_ln.log
timeout 300 $N -k regex:conv1d_igemm_pipe -s 60 -c 1 -o gpurun_out/r1n_conv_pipe python tools/vq_encode_bench.py > gpurun_out/r1n_ncu_conv.log 2>&1; tail -2 gpurun_out/r1n_ncu_conv.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r1n_launches_step.csv python bench.py --steps 2 --warmup 1 --profile-run --no-cpu-baseline --no-vq-encode --no-e2e > gpurun_out/r1n_ncu_bench.log 2>&1; python tools/summarize_launches.py gpurun_out/r1n_launches_step.csv 2>/dev/null | head -14
ls -la gpurun_out/*.ncu-rep
