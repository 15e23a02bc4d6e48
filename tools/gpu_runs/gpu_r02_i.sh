set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_vit_train_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/i_pytest.txt
tail -3 gpurun_out/i_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 20 > gpurun_out/i_train_n2.json 2> gpurun_out/i_train_n2.err
timeout 600 python bench.py --mode train --steps 20 > gpurun_out/i_train_n1.json 2> gpurun_out/i_train_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 > gpurun_out/i_bench_n2.json 2> gpurun_out/i_bench_n2.err
