mkdir -p gpurun_out
timeout 900 python tools/tile_sweep.py > gpurun_out/aa_tile_sweep.txt 2>&1
cat gpurun_out/aa_tile_sweep.txt
