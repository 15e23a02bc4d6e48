mkdir -p gpurun_out
timeout 600 python bench.py --mode train --steps 30 2> gpurun_out/z_train.err | grep "^{" > gpurun_out/z_train.json
python -c "import json;d=json.load(open('gpurun_out/z_train.json'));print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'])"
timeout 600 python bench.py --mode train --steps 30 --batch 64 2> gpurun_out/z_train.err | grep "^{" > gpurun_out/z_train64.json
python -c "import json;d=json.load(open('gpurun_out/z_train64.json'));print('train64', d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 300 python tools/train_timeline.py --brief > gpurun_out/z_tl.txt 2>&1
head -2 gpurun_out/z_tl.txt; grep " ms " gpurun_out/z_tl.txt | head -8
timeout 900 python bench.py --no-cpu-baseline 2> gpurun_out/z_bench.err | grep "^{" > gpurun_out/z_bench.json
python -c "
import json
d=json.load(open('gpurun_out/z_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'fp16',round(d['e2e_fp16_feed']['value']),'roof',d['roofline']['frac'], d['roofline']['achieved'], d['clocks']['sm_mhz'])
"
