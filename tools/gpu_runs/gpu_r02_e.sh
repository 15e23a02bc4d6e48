set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_vit_gpu.py tests/test_hub_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/e_pytest.txt
tail -3 gpurun_out/e_pytest.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/e_bench2.json 2>> gpurun_out/e_bench.err
