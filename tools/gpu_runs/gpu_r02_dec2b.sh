mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_codec_gpu.py tests/test_hub_gpu.py tests/test_vit_train_gpu.py -x -q > gpurun_out/dec2b_pytest.txt 2>&1
tail -3 gpurun_out/dec2b_pytest.txt
timeout 120 python tools/colsum_bench.py > gpurun_out/dec2b_colsum.txt 2>&1
cat gpurun_out/dec2b_colsum.txt
timeout 240 python tools/codec_bench.py --rows 32,128,512,2048,12096,131072,524288 --out gpurun_out/dec2b_codecbench.txt > gpurun_out/dec2b_codecbench.log 2>&1
tail -45 gpurun_out/dec2b_codecbench.log | cut -c1-100
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'rans_decode_lanes2' -c 1 -f -o gpurun_out/dec2b_ncu_lanes python tools/ncu_codec.py > /dev/null 2>&1
ncu -i gpurun_out/dec2b_ncu_lanes.ncu-rep --page raw --csv > gpurun_out/dec2b_ncu_lanes_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/dec2b_ncu_lanes_raw.csv > gpurun_out/dec2b_ncu_lanes_summary.txt 2>&1
grep -E "^==|duration|inst_executed.sum|issue_active|warps_active|registers|stalled" gpurun_out/dec2b_ncu_lanes_summary.txt | cut -c1-140
