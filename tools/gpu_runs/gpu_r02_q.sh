mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 20 > gpurun_out/q_train_n2.json 2> gpurun_out/q_train_n2.err
echo "rc=$?"
tail -c 1500 gpurun_out/q_train_n2.err
head -c 600 gpurun_out/q_train_n2.json
