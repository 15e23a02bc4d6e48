mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_hub_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -3 gpurun_out/r_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'fp16',round(d['e2e_fp16_feed']['value']),'raw',round(d['e2e_raw_uint8']['value']),'roof',d['roofline']['frac'], d['clocks'])
print(d['e2e'])
"
