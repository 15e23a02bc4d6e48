mkdir -p gpurun_out/f2
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/f2/multi_gpu_pytest.txt 2>&1
tail -3 gpurun_out/f2/multi_gpu_pytest.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 20 2> gpurun_out/f2/train_n2.err | grep "^{" > gpurun_out/f2/train_n2.json
python -c "
import json; d=json.loads(open('gpurun_out/f2/train_n2.json').read().strip().splitlines()[-1]); print('train n2', d['ms_per_step'], d['value'], d['collective'], d['clocks'])"
