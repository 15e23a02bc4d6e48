mkdir -p gpurun_out
S="compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 --launch-timeout 120"
for spec in "tests/test_codec_gpu.py::-k::not full_size" "tests/test_eb_gpu.py::-k::not large_rows" "tests/test_distortion_gpu.py::-k::golden or lossy or temperature" "tests/test_vit_train_gpu.py::-k::transpose or quickgelu or attention_backward" "tests/test_preprocess_gpu.py::-k::matches_pil"; do
f=${spec%%::*}; k=${spec##*::}
timeout 900 $S python -m pytest $f -m gpu -q -x -p no:cacheprovider -k "$k" > gpurun_out/race_tmp.txt 2>&1
echo "== $f -k '$k': rc=$? $(grep -c 'hazard detected' gpurun_out/race_tmp.txt) hazards"; tail -2 gpurun_out/race_tmp.txt
grep "hazard detected\|Thread" gpurun_out/race_tmp.txt | grep -v "at::native" | head -6
done
