mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --mode train --steps 30 2> gpurun_out/af_train.err | grep "^{" > gpurun_out/af_train.json
python -c "import json;d=json.load(open('gpurun_out/af_train.json'));print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
timeout 600 python bench.py --mode train --steps 30 --batch 64 2>> gpurun_out/af_train.err | grep "^{" > gpurun_out/af_train64.json
python -c "import json;d=json.load(open('gpurun_out/af_train64.json'));print('train64', d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/af_launches_train.csv python bench.py --mode train --steps 1 --warmup 3 > /dev/null 2>&1
timeout 300 python tools/train_timeline.py --brief > gpurun_out/af_tl.txt 2>&1
