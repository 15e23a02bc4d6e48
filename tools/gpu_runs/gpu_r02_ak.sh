mkdir -p gpurun_out
timeout 300 python tools/e2e_timeline.py 1512 1512 u8 > gpurun_out/ak_e2e_u8.txt 2>&1
cat gpurun_out/ak_e2e_u8.txt
timeout 300 python tools/e2e_timeline.py 1512 1512 > gpurun_out/ak_e2e_f16.txt 2>&1
cat gpurun_out/ak_e2e_f16.txt
