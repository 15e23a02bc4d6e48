mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)"; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done) > gpurun_out/n8_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --roofline-seconds 0.5 2> gpurun_out/n8_bench.err | grep "^{" > gpurun_out/n8_bench.json
LL_NO_NUMA=1 timeout 400 $TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --roofline-seconds 0.5 2> gpurun_out/n8_bench_nonuma.err | grep "^{" > gpurun_out/n8_bench_nonuma.json
timeout 400 $TR --master-port 29523 bench.py --gpus 8 --mode train --steps 20 2> gpurun_out/n8_train.err | grep "^{" > gpurun_out/n8_train.json
python - <<'PY'
import json
for f in ("n8_bench", "n8_bench_nonuma"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "u8", round(d["e2e_clip_uint8"]["value"]), "raw", round(d["e2e_raw_uint8"]["value"]), "link", d["e2e"].get("h2d_link_gbps"), d["e2e"].get("h2d_link_gbps_min_over_ranks"), d["e2e"].get("host_numa_binding"))
    except Exception as e:
        print(f, "failed", e)
try:
    d = json.load(open("gpurun_out/n8_train.json"))
    print("train", d["ms_per_step"], round(d["value"]), round(d["e2e"]["value"]), d["collective"])
except Exception as e:
    print("train failed", e)
PY
