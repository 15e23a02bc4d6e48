#!/bin/bash
# Staged GPU check: each stage under its own timeout so a hung kernel cannot eat the box.
# usage: tools/gpu_stage.sh [stage ...]   (default: all)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -s KILL "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
STAGES=${@:-codec gemm vit eb hub smoke bench}
for s in $STAGES; do
  case $s in
    codec) run codec 300 python -m pytest tests/test_codec_gpu.py -m gpu -x -q ;;
    gemm)  run gemm 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q ;;
    vit)   run vit 300 python -m pytest tests/test_vit_gpu.py -m gpu -q ;;
    eb)    run eb 300 python -m pytest tests/test_eb_gpu.py -m gpu -q ;;
    dist)  run dist 300 python -m pytest tests/test_distortion_gpu.py tests/test_training_gpu.py -m gpu -q ;;
    hyper) run hyper 300 python -m pytest tests/test_hyperprior.py -m gpu -q ;;
    prep)  run prep 300 python -m pytest tests/test_preprocess_gpu.py -m gpu -q ;;
    vittrain) run vittrain 600 python -m pytest tests/test_vit_train_gpu.py -m gpu -q -x ;;
    hub)   run hub 300 python -m pytest tests/test_hub_gpu.py -m gpu -q ;;
    all)   run all 900 python -m pytest tests -m gpu -x -q ;;
    smoke) run smoke 200 python __graft_entry__.py smoke ;;
    bench) run bench 400 python bench.py --steps 10 --warmup 3 ;;
    ab_ln) for i in 1 2; do
             run ab_sep$i 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fuse-ln
             run ab_fused$i 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
           done ;;
    ab_attn) for i in 0 2 0 2; do run ab_attn$i 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --attn-impl $i; done ;;
    e2eprobe) run e2eprobe 300 python tools/e2e_probe.py ;;
    codecbench) run codecbench 600 python tools/codec_bench.py --out gpurun_out/codecbench.txt ;;
    ncu_codec) run ncu_codec 600 ncu --set full --clock-control none --import-source on -k regex:lanes -s 17 -c 2 -f -o gpurun_out/prof_codec python tools/codec_bench.py --rows 131072
               ncu -i gpurun_out/prof_codec.ncu-rep --page raw --csv > gpurun_out/prof_codec_raw.csv 2>/dev/null
               ncu -i gpurun_out/prof_codec.ncu-rep --page source --csv > gpurun_out/prof_codec_source.csv 2>/dev/null ;;
    ncu_train) run ncu_train 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --only unfrozen --batches 256 --steps 1 --warmup 1 ;;
    sweep) run sweep 600 python tools/sweep.py --out gpurun_out/sweep.txt ;;
    trainbench) run trainbench 600 python tools/train_bench.py --out gpurun_out/trainbench.txt ;;
    trainbench_unfrozen) run trainbench_unfrozen 600 python tools/train_bench.py --only unfrozen --steps 10 --out gpurun_out/trainbench_unfrozen.txt ;;
    gemmbench) run gemmbench 200 python tools/gemm_bench.py 256 ;;
    pairbench) run pairbench0 200 env LL_PAIR=0 LL_GEMM_ONLY=1 python tools/gemm_bench.py 256
               run pairbench1 200 env LL_PAIR=1 LL_GEMM_ONLY=1 python tools/gemm_bench.py 256
               run pairbench1_512 200 env LL_PAIR=1 LL_GEMM_ONLY=1 python tools/gemm_bench.py 512
               run pairbench0_512 200 env LL_PAIR=0 LL_GEMM_ONLY=1 python tools/gemm_bench.py 512 ;;
    pair378) run pair378_bn256 200 env LL_PAIR=1 LL_BN=256 LL_GEMM_ONLY=1 python tools/gemm_bench.py 378
             run pair378_auto 200 env LL_PAIR=1 LL_GEMM_ONLY=1 python tools/gemm_bench.py 378
             run single378_auto 200 env LL_PAIR=0 LL_GEMM_ONLY=1 python tools/gemm_bench.py 378
             run pair756_bn256 200 env LL_PAIR=1 LL_BN=256 LL_GEMM_ONLY=1 python tools/gemm_bench.py 756 ;;
    epidbg) for d in 0 0x100 0x1100 0x300 0x700 0xf00 0x1f00; do run epidbg_$d 200 env LL_DBG=$d LL_PAIR=1 LL_BN=256 LL_GEMM_ONLY=1 python tools/gemm_bench.py 378; done ;;
    bnsweep) for bn in 128 192; do run pair378_bn$bn 200 env LL_PAIR=1 LL_BN=$bn LL_GEMM_ONLY=1 python tools/gemm_bench.py 378; done ;;
    ncu_vit) run ncu_vit 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 50 -c 10 -f -o gpurun_out/prof_vit_gemms python tools/vit_once.py
             ncu -i gpurun_out/prof_vit_gemms.ncu-rep --page raw --csv > gpurun_out/prof_vit_gemms_raw.csv 2>/dev/null
             python tools/ncu_gemm_traffic.py gpurun_out/prof_vit_gemms_raw.csv 1512 gpurun_out/ncu_vit_gemms_dram_traffic.json
             python tools/ncu_summary.py gpurun_out/prof_vit_gemms_raw.csv > gpurun_out/ncu_vit_gemms_summary.txt
             [ "$(stat -c %s gpurun_out/prof_vit_gemms.ncu-rep 2>/dev/null || echo 0)" -gt 40000000 ] && rm -f gpurun_out/prof_vit_gemms.ncu-rep ;;
    ncu_small) run ncu_small 600 ncu --set full --clock-control none --import-source on -k regex:"attention" -s 10 -c 2 -f -o gpurun_out/prof_small python tools/gemm_bench.py 378
               ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_attention_raw.csv 2>/dev/null ;;
    ncu_gemm) run ncu_gemm 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 5 -f -o gpurun_out/prof_gemm python tools/gemm_bench.py 256 ;;
    ncu_list) run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ;;
  esac
done
