mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1m_pytest.log 2>&1; tail -3 gpurun_out/r1m_pytest.log
for v in 4 5; do TTTS_ATTN_VER=$v timeout 120 python tools/attn_prof.py 2>&1 | sed "s/^/VER=$v /"; done | tee gpurun_out/r1m_attn_ab.txt
rm -f /tmp/kernels_ab.npz
(TTTS_STFT_V1=1 TTTS_VQ_V1=1 timeout 300 python tools/kernels_ab.py; timeout 300 python tools/kernels_ab.py) 2>&1 | tee gpurun_out/r1m_kernels_ab.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vq-encode > gpurun_out/r1m_bench_quick.json 2> gpurun_out/r1m_bench_quick.err; cut -c1-400 gpurun_out/r1m_bench_quick.json
