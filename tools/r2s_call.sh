# Round 2, GPU call S: where does a KV-cache decode step spend its 1.6 ms?  Launch list of tools/decode_bench.py (B = 1).
mkdir -p gpurun_out
T=r2s
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dec_ -s 2000 -c 4000 --csv --log-file gpurun_out/${T}_launches_decode_b1.csv python tools/decode_bench.py 1 32 64 > gpurun_out/${T}_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches_decode_b1.csv > gpurun_out/${T}_launches_decode_b1_summary.txt 2>&1; head -n 14 gpurun_out/${T}_launches_decode_b1_summary.txt
rm -f gpurun_out/${T}_launches_decode_b1.csv
timeout 300 python tools/decode_bench.py 1 32 256 > gpurun_out/${T}_decode_b1.json 2>&1; cut -c1-900 gpurun_out/${T}_decode_b1.json
