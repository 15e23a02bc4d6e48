mkdir -p gpurun_out
timeout 300 python tools/train_timeline.py --brief > gpurun_out/o_tl.txt 2>&1
sed -n 1,12p gpurun_out/o_tl.txt;  grep -n "inside:" gpurun_out/o_tl.txt | head
timeout 600 python bench.py --mode train --steps 30 > gpurun_out/o_train.json 2> gpurun_out/o_train.err
python -c "import json;d=json.load(open('gpurun_out/o_train.json'));print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'])"
