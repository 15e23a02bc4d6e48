set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/a_pytest.txt
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
timeout 300 python tools/torch_vit_yardstick.py --out gpurun_out/a_yardstick.txt > /dev/null 2> gpurun_out/a_yardstick.err
timeout 600 python bench.py --mode train --steps 10 > gpurun_out/a_train.json 2> gpurun_out/a_train.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err
tail -5 gpurun_out/a_pytest.txt
