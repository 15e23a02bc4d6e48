mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/p_pytest.txt
tail -4 gpurun_out/p_pytest.txt
timeout 600 python bench.py --mode train --steps 30 --batch 64 > gpurun_out/p_train_b64.json 2> gpurun_out/p_train.err
python -c "import json;d=json.load(open('gpurun_out/p_train_b64.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'])"
