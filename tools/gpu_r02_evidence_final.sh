# Round-2 final evidence run (one gpurun call, one GPU): bench lines, sweeps, ncu launch lists and --set full captures.
set -x
mkdir -p gpurun_out/evf
O=gpurun_out/evf
timeout 400 python bench.py 2> $O/bench_n1.err | grep "^{" > $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> $O/bench_ref.err | grep "^{" > $O/bench_reference_arm.json
timeout 300 python tools/torch_vit_yardstick.py --out $O/torch_vit_yardstick.txt > /dev/null 2> $O/yard.err
timeout 600 python tools/sweep.py --out $O/sweep.txt > /dev/null 2> $O/sweep.err
timeout 600 python tools/codec_bench.py --out $O/codecbench.txt > /dev/null 2> $O/codec.err
timeout 600 python bench.py --mode train --steps 30 2> $O/train_n1.err | grep "^{" > $O/train_n1.json
timeout 600 python bench.py --mode train --steps 30 --frozen --distortion lossy_Z 2>> $O/train_n1.err | grep "^{" > $O/train_frozen_lossyz_n1.json
timeout 600 python bench.py --mode train --steps 30 --batch 64 2>> $O/train_n1.err | grep "^{" > $O/train_b64_n1.json
timeout 300 python tools/train_timeline.py --brief > $O/train_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_encode.csv python tools/ncu_one_pass.py 1512 2 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -s 50 -c 10 -f -o $O/ncu_vit_gemms python tools/ncu_one_pass.py 1512 2 > /dev/null 2>&1
ncu -i $O/ncu_vit_gemms.ncu-rep --page raw --csv > $O/ncu_vit_gemms_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/ncu_vit_gemms_raw.csv > $O/ncu_vit_gemms_summary.txt 2>&1
rm -f $O/ncu_vit_gemms.ncu-rep
ls -la $O
