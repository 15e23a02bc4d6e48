mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_gpu.py tests/test_vit_train_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for n in 32 64 128 256 512 1512; do
LL_GEMM_ONLY=1 timeout 300 python tools/gemm_bench.py $n 2>&1 | grep -v "^$"
done > gpurun_out/y_gemm_small2.txt
grep "per pass" gpurun_out/y_gemm_small2.txt
timeout 600 python tools/ab_cls_last.py 2>&1 | cut -c1-60
