mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_train_gpu.py tests/test_training_gpu.py tests/test_ref_pinned_gpu.py tests/test_native_abi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --mode train --steps 30 2> gpurun_out/ab_train.err | grep "^{" > gpurun_out/ab_train.json
python -c "import json;d=json.load(open('gpurun_out/ab_train.json'));print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'])"
