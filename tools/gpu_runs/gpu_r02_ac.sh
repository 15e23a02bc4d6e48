mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_train_gpu.py tests/test_training_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --mode train --steps 30 2> gpurun_out/ac_train.err | grep "^{" > gpurun_out/ac_train.json
python -c "import json;d=json.load(open('gpurun_out/ac_train.json'));print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'])"
timeout 300 python tools/train_timeline.py --brief > gpurun_out/ac_tl.txt 2>&1
head -2 gpurun_out/ac_tl.txt; grep " ms " gpurun_out/ac_tl.txt | head -14
