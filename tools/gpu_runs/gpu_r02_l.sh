set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/l_pytest.txt
tail -3 gpurun_out/l_pytest.txt
timeout 600 python bench.py --mode train --steps 30 > gpurun_out/l_train.json 2> gpurun_out/l_train.err
timeout 600 python bench.py --mode train --steps 30 --batch 64 > gpurun_out/l_train_b64.json 2>> gpurun_out/l_train.err
timeout 300 python tools/train_timeline.py > gpurun_out/l_timeline.txt 2>&1
timeout 300 python tools/sweep.py --out gpurun_out/l_sweep.txt > /dev/null 2>&1
tail -3 gpurun_out/l_sweep.txt
