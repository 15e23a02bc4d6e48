mkdir -p gpurun_out
S="compute-sanitizer --tool synccheck --error-exitcode 99 --launch-timeout 120"
for spec in "tests/test_codec_gpu.py::-k::not full_size" "tests/test_vit_train_gpu.py::-k::layernorm_backward or attention_backward or transpose" "tests/test_gemm_gpu.py::-k::dual_output or splitk or gemm_nn"; do
f=${spec%%::*}; k=${spec##*::}
timeout 900 $S python -m pytest $f -m gpu -q -x -p no:cacheprovider -k "$k" > gpurun_out/sync_tmp.txt 2>&1
echo "== $f -k '$k': rc=$?"; tail -2 gpurun_out/sync_tmp.txt
grep -i "barrier error\|divergent\|Invalid" gpurun_out/sync_tmp.txt | head -5
done
