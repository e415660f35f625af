# Round 2, GPU call AN: `ncu --set full` of the layout kernels and the 16-byte-load bias gradient (HBM-bound kernels of the GEMM route).
mkdir -p gpurun_out
T=r2an
ITERS=1 timeout 55 ncu --set full --clock-control none --import-source on -k regex:"cl_split|cl_unpack|bgrad4" -s 2 -c 4 -o gpurun_out/${T}_layout -f python tools/gemm_conv_prof.py > gpurun_out/${T}_ncu.log 2>&1
timeout 20 python tools/ncu_summary.py gpurun_out/${T}_layout.ncu-rep > gpurun_out/${T}_layout_ncu_full.txt 2>&1; grep -E "Kernel Name|dram__bytes_(read|write).sum \[|gpu__time_duration" gpurun_out/${T}_layout_ncu_full.txt | cut -c1-260
