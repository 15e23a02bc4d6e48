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
    hub)   run hub 300 python -m pytest tests/test_hub_gpu.py -m gpu -q ;;
    all)   run all 900 python -m pytest tests -m gpu -x -q ;;
    smoke) run smoke 200 python __graft_entry__.py smoke ;;
    bench) run bench 400 python bench.py --steps 10 --warmup 3 ;;
  esac
done
