mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_train_gpu.py tests/test_vit_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for m in 3 7 3 7; do
timeout 300 python tools/train_timeline.py --brief --fusion $m > gpurun_out/ae_tl_$m.txt 2>&1
echo "fusion=$m"; sed -n 2p gpurun_out/ae_tl_$m.txt; grep "gemm_tcgen05_kernel<__nv_bfloat16, 256, 2\|quickgelu_forward" gpurun_out/ae_tl_$m.txt | grep " ms "
done
