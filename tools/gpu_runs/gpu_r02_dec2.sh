mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_codec_gpu.py tests/test_hub_gpu.py -x -q 2>&1 | tail -5
timeout 240 python tools/codec_bench.py --rows 512,2048,12096,32768,131072,524288 --out gpurun_out/dec2_codecbench.txt > gpurun_out/dec2_codecbench.log 2>&1
tail -40 gpurun_out/dec2_codecbench.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'lanes' -c 3 -f -o gpurun_out/dec2_ncu_lanes python tools/ncu_codec.py > /dev/null 2>&1
ncu -i gpurun_out/dec2_ncu_lanes.ncu-rep --page raw --csv > gpurun_out/dec2_ncu_lanes_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/dec2_ncu_lanes_raw.csv > gpurun_out/dec2_ncu_lanes_summary.txt 2>&1
grep -E "^==|duration|inst_executed.sum|issue_active|warps_active|registers" gpurun_out/dec2_ncu_lanes_summary.txt | tail -40
