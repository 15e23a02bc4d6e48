mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/n8t_pytest.txt
cat gpurun_out/n8t_pytest.txt
