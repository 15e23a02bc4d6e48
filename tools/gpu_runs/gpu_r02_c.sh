set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_vit_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/c_pytest1.txt
tail -3 gpurun_out/c_pytest1.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/c_pytest.txt
tail -3 gpurun_out/c_pytest.txt
timeout 300 python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
