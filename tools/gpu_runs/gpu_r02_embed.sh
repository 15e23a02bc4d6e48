mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_vit_train_gpu.py tests/test_training_gpu.py -x -q > gpurun_out/embed_pytest.txt 2>&1
tail -3 gpurun_out/embed_pytest.txt
timeout 300 python bench.py --mode train --steps 30 2> gpurun_out/embed_train.err | grep "^{" > gpurun_out/embed_train.json
python -c "
import json; d=json.loads(open('gpurun_out/embed_train.json').read().strip().splitlines()[-1]); print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['loss'])"
