set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_train_launches.csv python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/g_train.log 2>&1
tail -2 gpurun_out/g_train.log
