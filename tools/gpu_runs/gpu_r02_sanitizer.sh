mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 120"
(
echo "# compute-sanitizer --tool memcheck over the GPU tests of the kernels added / rewritten in the second half of round 2"
timeout 700 $S python -m pytest tests/test_codec_gpu.py -m gpu -q -x -p no:cacheprovider -k "stream_per_lane" 2>&1 | tail -4
echo "rc=$?"
timeout 700 $S python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider -k "splitk or dual_output or gemm_nn or quickgelu_backward" 2>&1 | tail -4
echo "rc=$?"
timeout 700 $S python -m pytest tests/test_vit_train_gpu.py tests/test_vit_gpu.py -m gpu -q -x -p no:cacheprovider -k "cls_only or fusions or layernorm_backward or forward_views" 2>&1 | tail -4
echo "rc=$?"
) > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
