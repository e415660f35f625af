# Round 2, GPU call Q: ncu --set full captures of the kernels that dominate the diffusion / VQ-VAE-GAN steps after the rework
# (attn_bias forward / dq / dkv, conv1d_wgrad2, the split-bf16 GEMM inside the diffusion step), launch list of the tensor-core diffusion step.
mkdir -p gpurun_out
T=r2q
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:attn_bias_fwd_kernel -s 12 -c 1 -o gpurun_out/${T}_attn_bias_fwd -f python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_ncu1.log 2>&1; tail -n 1 gpurun_out/${T}_ncu1.log | cut -c1-200
timeout 400 $N -k regex:attn_bias_bwd_dq_kernel -s 12 -c 1 -o gpurun_out/${T}_attn_bias_bwd_dq -f python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_ncu2.log 2>&1; tail -n 1 gpurun_out/${T}_ncu2.log | cut -c1-200
timeout 400 $N -k regex:attn_bias_bwd_dkv_kernel -s 12 -c 1 -o gpurun_out/${T}_attn_bias_bwd_dkv -f python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_ncu3.log 2>&1; tail -n 1 gpurun_out/${T}_ncu3.log | cut -c1-200
TTTS_DIFF_TC=0 timeout 400 $N -k regex:conv1d_wgrad2_kernel -s 40 -c 1 -o gpurun_out/${T}_wgrad2 -f python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_ncu4.log 2>&1; tail -n 1 gpurun_out/${T}_ncu4.log | cut -c1-200
timeout 400 $N -k regex:gemm2_bf16_kernel -s 60 -c 2 -o gpurun_out/${T}_diff_gemm -f python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_ncu5.log 2>&1; tail -n 1 gpurun_out/${T}_ncu5.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches_diffusion_tc_b32.csv python tools/diffusion_step_bench.py 32 1 > gpurun_out/${T}_ncu6.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_diffusion_tc_b32.csv > gpurun_out/${T}_launches_diffusion_tc_b32_summary.txt 2>&1; head -n 24 gpurun_out/${T}_launches_diffusion_tc_b32_summary.txt
rm -f gpurun_out/${T}_launches_diffusion_tc_b32.csv
ls -la gpurun_out/${T}_*.ncu-rep
