set -x
mkdir -p gpurun_out
# launch list of one warm encode step (second pass): every launch with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches.csv python tools/ncu_one_pass.py 1512 2 > gpurun_out/d_launches.log 2>&1
# full capture of block 1's four GEMMs in the second pass (GEMM launch index: 50 of pass 1 + patch + 4 of block 0 = 55)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -s 55 -c 4 -f -o gpurun_out/d_gemm python tools/ncu_one_pass.py 1512 2 > gpurun_out/d_gemm.log 2>&1
ls -la gpurun_out/
