mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python tools/sweep.py --out gpurun_out/v_sweep.txt > /dev/null 2>&1
grep "beta5e-02" gpurun_out/v_sweep.txt | head -9
