mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_hub_gpu.py tests/test_ref_pinned_gpu.py tests/test_hyperprior.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/codec_bench.py --out gpurun_out/s_codecbench.txt > /dev/null 2>&1
grep "lane" gpurun_out/s_codecbench.txt | grep "131072\|524288\|32768"
