set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_vit_train_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/m_pytest.txt
tail -5 gpurun_out/m_pytest.txt
timeout 300 python tools/train_timeline.py --brief --splitk 0 > gpurun_out/m_tl_split0.txt 2>&1
timeout 300 python tools/train_timeline.py --brief --splitk 1 > gpurun_out/m_tl_split1.txt 2>&1
head -3 gpurun_out/m_tl_split0.txt gpurun_out/m_tl_split1.txt
