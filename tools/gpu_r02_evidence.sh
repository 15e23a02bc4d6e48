# Round-2 evidence run (one gpurun call): bench lines, yard-sticks, sweeps, ncu launch lists and --set full captures.
set -x
mkdir -p gpurun_out/ev
O=gpurun_out/ev
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
timeout 300 python tools/torch_vit_yardstick.py --out $O/torch_vit_yardstick.txt > /dev/null 2> $O/yard.err
timeout 600 python tools/sweep.py --out $O/sweep.txt > /dev/null 2> $O/sweep.err
timeout 600 python tools/codec_bench.py --out $O/codecbench.txt > /dev/null 2> $O/codec.err
timeout 600 python bench.py --mode train --steps 20 > $O/train_n1.json 2> $O/train_n1.err
timeout 600 python bench.py --mode train --steps 20 --frozen --distortion lossy_Z > $O/train_frozen_lossyz_n1.json 2>> $O/train_n1.err
timeout 600 python bench.py --mode train --steps 20 --batch 64 > $O/train_b64_n1.json 2>> $O/train_n1.err
timeout 300 python tools/train_timeline.py > $O/train_timeline.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_encode.csv python tools/ncu_one_pass.py 1512 2 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -s 50 -c 10 -f -o $O/ncu_vit_gemms python tools/ncu_one_pass.py 1512 2 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_tma -s 12 -c 1 -f -o $O/ncu_attention python tools/ncu_one_pass.py 1512 2 > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eb_forward|eb_backward|lanes_kernel|eb_rans_encode_kernel|gc_forward' -s 7 -c 8 -f -o $O/ncu_entropy python tools/ncu_entropy.py > /dev/null 2>&1
ls -la $O
