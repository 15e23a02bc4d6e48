mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/n8f_pytest.txt
cat gpurun_out/n8f_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --roofline-seconds 0.5 2> gpurun_out/n8f_bench.err | grep "^{" > gpurun_out/n8f_bench.json
timeout 400 $TR --master-port 29533 bench.py --gpus 8 --mode train --steps 20 2> gpurun_out/n8f_train.err | grep "^{" > gpurun_out/n8f_train.json
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR2 --master-port 29535 bench.py --gpus 2 --mode train --steps 20 2> gpurun_out/n2f_train.err | grep "^{" > gpurun_out/n2f_train.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/n8f_bench.json"))
    print("encode n8 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "fp16", round(d["e2e_fp16_feed"]["value"]), "raw", round(d["e2e_raw_uint8"]["value"]), "link", d["e2e"].get("h2d_link_gbps_min_over_ranks"))
except Exception as e:
    print("bench failed", e)
for f in ("n8f_train", "n2f_train"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["ms_per_step"], round(d["value"]), round(d["e2e"]["value"]), d["collective"])
    except Exception as e:
        print(f, "failed", e)
PY
