mkdir -p gpurun_out
for m in 1 3 1 3; do
timeout 300 python tools/train_timeline.py --brief --fusion $m > gpurun_out/ad_tl_$m.txt 2>&1
echo "fusion=$m"; head -2 gpurun_out/ad_tl_$m.txt | tail -1; grep "gemm_tcgen05_kernel<__nv_bfloat16, 256, 2>\|transpose_pad" gpurun_out/ad_tl_$m.txt
done
