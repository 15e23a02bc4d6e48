mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
