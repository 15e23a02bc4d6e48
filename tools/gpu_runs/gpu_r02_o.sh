mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_vit_train_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8
timeout 600 python bench.py --mode train --steps 30 > gpurun_out/o_train.json 2> gpurun_out/o_train.err
python -c "import json;d=json.load(open('gpurun_out/o_train.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'])"
tail -3 gpurun_out/o_train.err
