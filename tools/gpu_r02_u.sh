mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eb_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/sweep.py --out gpurun_out/u_sweep.txt > /dev/null 2>&1
tail -2 gpurun_out/u_sweep.txt
