# Round 2, GPU call N: pipelined weight-gradient kernel (conv1d_wgrad2) + stride-1 input gradients through the forward implicit GEMM:
# parity of everything that differentiates a convolution (diffusion + VQ-VAE tapes), A/B of the diffusion and VQ-VAE-GAN steps.
mkdir -p gpurun_out
T=r2n
timeout 900 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_encoder.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -30
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  TFLOP/s %.1f  launches %d" % (d["ms_per_step"], d["step_tflops"], d["gpu_launches_per_step"]), {k: round(v["ms_per_step"],1) for k,v in d.get("roofline",{}).get("kernels",{}).items()})'
timeout 600 python tools/diffusion_step_bench.py 32 2 > gpurun_out/${T}_diffusion_new.json 2> gpurun_out/${T}_diffusion_new.err; python -c "$P" gpurun_out/${T}_diffusion_new.json
TTTS_WGRAD_V1=1 TTTS_DGRAD_FWD=0 timeout 600 python tools/diffusion_step_bench.py 32 2 > gpurun_out/${T}_diffusion_old.json 2> gpurun_out/${T}_diffusion_old.err; python -c "$P" gpurun_out/${T}_diffusion_old.json
TTTS_WGRAD_V1=0 TTTS_DGRAD_FWD=0 timeout 600 python tools/diffusion_step_bench.py 32 2 > gpurun_out/${T}_diffusion_wgrad2only.json 2> gpurun_out/${T}_diffusion_wgrad2only.err; python -c "$P" gpurun_out/${T}_diffusion_wgrad2only.json
timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_new.json 2> gpurun_out/${T}_vqvae_new.err; python -c "$P" gpurun_out/${T}_vqvae_new.json
TTTS_WGRAD_V1=1 TTTS_DGRAD_FWD=0 timeout 600 python tools/vqvae_step_bench.py 64 2 > gpurun_out/${T}_vqvae_old.json 2> gpurun_out/${T}_vqvae_old.err; python -c "$P" gpurun_out/${T}_vqvae_old.json
tail -2 gpurun_out/${T}_*.err | cut -c1-300
