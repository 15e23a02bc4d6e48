set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_train_gpu.py tests/test_training_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/f_pytest.txt
tail -5 gpurun_out/f_pytest.txt
timeout 600 python bench.py --mode train --steps 10 > gpurun_out/f_train.json 2> gpurun_out/f_train.err
LL_FLAT_OPTIMIZER=0 timeout 600 python bench.py --mode train --steps 10 > gpurun_out/f_train_noflat.json 2> gpurun_out/f_train_noflat.err
