"""ncu raw page of the encoder's first ten GEMM launches -> profiles/*_dram_traffic.json (per-launch DRAM bytes,
duration and algorithmic bytes; bench.py reports the block average as roofline.traffic).
usage: python tools/ncu_gemm_traffic.py prof_vit_gemms_raw.csv images out.json"""
import csv
import json
import sys


def main():
    raw, images, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    csv.field_size_limit(10 ** 9)
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    units = rows[1]
    M, Mp = 50 * images, 49 * images
    W = 768
    # (name, M, N, K, residual): launch order of ll_vit_forward after the patch-embed GEMM
    block = [("qkv(+ln_1)", M, 3 * W, W, 0), ("out_proj(+res,+stats)", M, W, W, 1), ("fc1(+ln_2,+gelu)", M, 4 * W, W, 0),
             ("fc2(+res,+stats)", M, W, 4 * W, 1)]
    order = [("patch_embed", Mp, W, 3 * 32 * 32, 0)] + [block[i % 4] for i in range(len(rows) - 3)]

    def to(v, u, want):
        v = float(v.replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
        return v * scale[u] / (1e6 if want == "MB" else 1)

    launches = []
    for r, (name, m, n, k, res) in zip(rows[2:], order):
        algo = 2 * (m * k + n * k + m * n * (1 + res)) + 4 * n
        launches.append(dict(gemm=name, M=m, N=n, K=k,
                             duration_us=to(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]], "us"),
                             dram_read_MB=to(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]], "MB"),
                             dram_write_MB=to(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]], "MB"),
                             algorithmic_MB=algo / 1e6))
    blk = launches[1:5]
    avg = sum(l["dram_read_MB"] + l["dram_write_MB"] for l in blk) / len(blk) * 1e6
    json.dump({
        "source": f"ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 50 -c 10 python tools/vit_once.py {images} "
                  f"({images}-image pass, fused LayerNorm build); cold caches per replay",
        "launches": launches,
        "per_launch_dram_bytes_layer_average": avg,
        "per_launch_algorithmic_bytes_layer_average": sum(l["algorithmic_MB"] for l in blk) / len(blk) * 1e6,
        "note": "per-launch figures; DRAM traffic at or below the algorithmic bytes means no wasted re-reads "
                "(part of D and of the next A operand stay in the 126 MB L2)",
    }, open(out, "w"), indent=1)
    print(f"block average: {avg / 1e6:.1f} MB DRAM per launch")


if __name__ == "__main__":
    main()
