mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_train_gpu.py tests/test_training_gpu.py tests/test_ref_pinned_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
for c in 1 0 1 0; do
LL_CLS_LAST_BLOCK=$c timeout 600 python bench.py --mode train --steps 30 2> gpurun_out/x_train.err | grep "^{" > gpurun_out/x_train_cls$c.json
python -c "import json;d=json.load(open('gpurun_out/x_train_cls$c.json'));print('cls=$c', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['loss'])"
done
