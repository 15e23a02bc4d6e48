mkdir -p gpurun_out
timeout 22 compute-sanitizer --tool memcheck --error-exitcode 99 python tools/race_bwd_kernels.py > gpurun_out/mem3_bwd.txt 2>&1
echo "bwd rc=$?"; tail -2 gpurun_out/mem3_bwd.txt
timeout 25 compute-sanitizer --tool memcheck --error-exitcode 99 python tools/ncu_codec.py 2085 > gpurun_out/mem3_codec.txt 2>&1
echo "codec rc=$?"; tail -2 gpurun_out/mem3_codec.txt
