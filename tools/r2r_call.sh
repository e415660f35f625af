# Round 2, GPU call R: attn_bias kernels re-tiled (64 threads, 8 x 8 register tiles): parity, per-launch times against r2q, step time.
mkdir -p gpurun_out
T=r2r
timeout 900 python -m pytest tests/test_gpu_diffusion.py -m gpu -q -rf > gpurun_out/${T}_pytest.log 2>&1
echo "== pytest rc=$?"; tail -n 4 gpurun_out/${T}_pytest.log | cut -c1-400
grep -h "^FAILED\|^ERROR\|^E  " gpurun_out/${T}_pytest.log | cut -c1-300 | head -20
P='import json,sys; d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], "ms/step %.1f  TFLOP/s %.1f  launches %d" % (d["ms_per_step"], d["step_tflops"], d["gpu_launches_per_step"]), d.get("losses", d.get("loss")))'
timeout 600 python tools/diffusion_step_bench.py 32 3 > gpurun_out/${T}_diffusion.json 2> gpurun_out/${T}_diffusion.err; python -c "$P" gpurun_out/${T}_diffusion.json; tail -n 3 gpurun_out/${T}_diffusion.err | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_bias -c 400 --csv --log-file gpurun_out/${T}_launches_attn.csv python tools/diffusion_step_bench.py 32 1 > gpurun_out/${T}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_attn.csv > gpurun_out/${T}_launches_attn_summary.txt 2>&1; head -n 14 gpurun_out/${T}_launches_attn_summary.txt
rm -f gpurun_out/${T}_launches_attn.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bias_fwd_kernel -s 12 -c 1 -o gpurun_out/${T}_attn_bias_fwd -f python tools/diffusion_step_bench.py 8 1 > gpurun_out/${T}_ncu1.log 2>&1; tail -n 1 gpurun_out/${T}_ncu1.log | cut -c1-200
