mkdir -p gpurun_out
S="compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 --launch-timeout 60"
for spec in "tests/test_vit_train_gpu.py::-k::layernorm_backward or embed_tokens or colsum" "tests/test_codec_gpu.py::-k::stream_per_lane_edge or quotient_paths or (stream_per_lane_kernels and beta5e-02)"; do
f=${spec%%::*}; k=${spec##*::}
timeout 55 $S python -m pytest $f -m gpu -q -x -p no:cacheprovider -k "$k" > gpurun_out/race3_tmp.txt 2>&1
echo "== $f -k '$k': rc=$? $(grep -c 'hazard detected' gpurun_out/race3_tmp.txt) hazards"; tail -2 gpurun_out/race3_tmp.txt
grep "hazard detected\|Thread" gpurun_out/race3_tmp.txt | grep -v "at::native" | head -6
cat gpurun_out/race3_tmp.txt >> gpurun_out/race3_all.txt
done
