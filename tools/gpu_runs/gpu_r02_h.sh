set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_train_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/h_pytest.txt
tail -3 gpurun_out/h_pytest.txt
for i in 1 2; do
timeout 600 python bench.py --mode train --steps 30 > gpurun_out/h_train_native_$i.json 2> gpurun_out/h_train.err
LL_NATIVE_BLOCKS=0 timeout 600 python bench.py --mode train --steps 30 > gpurun_out/h_train_python_$i.json 2>> gpurun_out/h_train.err
done
