mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'layernorm_backward' -c 1 -f -o gpurun_out/lnb_ncu python tools/bwd_kernels_bench.py > /dev/null 2>&1
ncu -i gpurun_out/lnb_ncu.ncu-rep --page raw --csv > gpurun_out/lnb_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/lnb_ncu_raw.csv | cut -c1-140
