mkdir -p gpurun_out/f2
O=gpurun_out/f2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 > /dev/null 2>&1
wc -l $O/launches_train.csv
python tools/launch_summary.py $O/launches_train.csv 2>&1 | head -40
