mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_vit_train_gpu.py -x -q > gpurun_out/lnb_pytest.txt 2>&1
tail -3 gpurun_out/lnb_pytest.txt
timeout 120 python tools/bwd_kernels_bench.py > gpurun_out/lnb_bwd_kernels.txt 2>&1
cat gpurun_out/lnb_bwd_kernels.txt
timeout 300 python bench.py --mode train --steps 30 2> gpurun_out/lnb_train.err | grep "^{" > gpurun_out/lnb_train.json
python -c "
import json; d=json.loads(open('gpurun_out/lnb_train.json').read().strip().splitlines()[-1]); print('train', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
