mkdir -p gpurun_out
timeout 50 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 --launch-timeout 60 python tools/race_bwd_kernels.py > gpurun_out/race4.txt 2>&1
echo "rc=$? $(grep -c 'hazard detected' gpurun_out/race4.txt) hazards"; tail -3 gpurun_out/race4.txt
