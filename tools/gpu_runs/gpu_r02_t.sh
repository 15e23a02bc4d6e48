mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'lanes_kernel' -c 2 -f -o gpurun_out/t_ncu_lanes python tools/ncu_entropy.py > /dev/null 2>&1
ncu -i gpurun_out/t_ncu_lanes.ncu-rep --page raw --csv > gpurun_out/t_ncu_lanes_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/t_ncu_lanes_raw.csv > gpurun_out/t_ncu_lanes_summary.txt 2>&1
head -80 gpurun_out/t_ncu_lanes_summary.txt
