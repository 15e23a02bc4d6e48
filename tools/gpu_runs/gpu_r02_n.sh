set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_train_gpu.py tests/test_training_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/n_pytest.txt
tail -5 gpurun_out/n_pytest.txt
timeout 300 python tools/train_timeline.py --brief > gpurun_out/n_tl.txt 2>&1
head -3 gpurun_out/n_tl.txt; grep " ms " gpurun_out/n_tl.txt | head -16
