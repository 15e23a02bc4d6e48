set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hyperprior.py tests/test_ref_pinned_gpu.py tests/test_vit_train_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/j_pytest.txt
tail -3 gpurun_out/j_pytest.txt
