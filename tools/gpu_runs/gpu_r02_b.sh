set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/b_pytest.txt
timeout 600 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
timeout 600 python tools/sweep.py --out gpurun_out/b_sweep.txt > /dev/null 2> gpurun_out/b_sweep.err
timeout 600 python bench.py --mode train --steps 10 > gpurun_out/b_train.json 2> gpurun_out/b_train.err
tail -5 gpurun_out/b_pytest.txt
