mkdir -p gpurun_out
S="compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 --launch-timeout 120"
timeout 900 $S python -m pytest tests/test_codec_gpu.py -m gpu -q -x -p no:cacheprovider -k "stream_per_lane" > gpurun_out/race_codec.txt 2>&1
echo "rc=$?"; tail -3 gpurun_out/race_codec.txt; grep -c "hazard detected" gpurun_out/race_codec.txt
timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_hub_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
