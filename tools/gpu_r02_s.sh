mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/codec_bench.py --out gpurun_out/s_codecbench.txt > /dev/null 2>&1
grep "lane" gpurun_out/s_codecbench.txt | grep "131072\|524288\|32768"
timeout 300 python tools/sweep.py --out gpurun_out/s_sweep.txt > /dev/null 2>&1
tail -3 gpurun_out/s_sweep.txt
