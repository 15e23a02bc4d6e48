mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 120"
for spec in "tests/test_vit_gpu.py::-k::cls_only or two_layers or b32_vs_golden" "tests/test_hub_gpu.py::-k::roundtrip or file_format or stream_matches" "tests/test_training_gpu.py::-k::prefetches or clip_step"; do
f=${spec%%::*}; k=${spec##*::}
timeout 1200 $S python -m pytest $f -m gpu -q -x -p no:cacheprovider -k "$k" > gpurun_out/mem_tmp.txt 2>&1
echo "== $f -k '$k': rc=$?"; tail -2 gpurun_out/mem_tmp.txt
grep -i "Invalid\|out of bounds\|misaligned" gpurun_out/mem_tmp.txt | head -5
done
