# Round-2 closing run after the decoder / LayerNorm-backward / column-sum / token-assembly changes: the whole GPU suite, smoke(),
# the default bench line, the training bench and its timeline.
mkdir -p gpurun_out/f2
O=gpurun_out/f2
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1
tail -3 $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py 2> $O/bench_n1.err | grep "^{" > $O/bench_n1.json
python -c "
import json; d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['roofline']['frac'], d['decode']['value'], d['decode']['ms_per_call'], d['decode']['e2e']['value'], d['clocks'])"
timeout 300 python bench.py --mode train --steps 30 2> $O/train_n1.err | grep "^{" > $O/train_n1.json
python -c "
import json; d=json.loads(open('$O/train_n1.json').read().strip().splitlines()[-1]); print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
timeout 300 python tools/train_timeline.py --brief > $O/train_timeline.txt 2>&1
sed -n 1,8p $O/train_timeline.txt
grep -A30 "cudaDeviceSynchronize at" $O/train_timeline.txt | grep " ms " | head -24
