"""Codec kernels alone: EB quantise + rANS encode + pack, and decode, stream-per-warp vs stream-per-lane
kernels over a range of row counts (SURVEY 8d "honest caveat": the HBM figure of the entropy kernels is
shown on large row counts).  Algorithmic bytes: encode = 2 B x 512 (fp16 z) + stream bytes + 4 B length per
row; decode = stream bytes + 4 B x 512 (fp32 z_hat) per row.

    python tools/codec_bench.py [--out file]
"""
import argparse
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore", message="no CLIP weights given")
import hubconf  # noqa: E402
from lossyless_b200 import _native as N  # noqa: E402


def timed(fn, iters, reps=3):
    """Best of `reps` event-timed loops (the first loops of a new size also pay the caching allocator's
    cudaMalloc of the scratch / output buffers)."""
    for _ in range(3):
        fn()
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--rows", default="512,2048,8192,32768,131072,524288")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6538.0))
    lines = [f"# codec kernels on {torch.cuda.get_device_name(0)}; z ~ N(0, 0.45^2) fp16 [rows, 512]; HBM peak {hbm:.0f} GB/s "
             f"({'measured' if peaks else 'fallback'}); times are CUDA-event averages incl. buffer allocation and the pack launches",
             f"{'ckpt':10s} {'rows':>7s} {'kernel':>6s} {'enc+pack us':>12s} {'GB/s':>8s} {'%HBM':>6s} {'decode us':>10s} {'GB/s':>8s} {'%HBM':>6s} {'bits/img':>9s}"]
    g = torch.Generator(device=dev).manual_seed(0)
    for tag, entry in (("beta5e-02", hubconf.clip_compressor_b005), ("beta1e-02", hubconf.clip_compressor_b001)):
        comp, _ = entry(device=dev)
        eb = comp.entropy_bottleneck
        es = comp._exp_scale()
        for rows in [int(r) for r in args.rows.split(",")]:
            z = (torch.randn(rows, 512, generator=g, device=dev) * 0.45).half()
            ref = None
            zref = None
            for impl, name in ((1, "warp"), (2, "lane")):
                N.call("ll_rans_set_impl", impl)
                iters = 20 if rows <= 32768 else 5
                ms_e, (packed, offs, lens, _) = timed(lambda: eb.compress_to_device(z, bias=comp.biasing, exp_scale=es, mode=0), iters)
                nbytes = int(lens.sum())
                if ref is None:
                    ref = packed[:nbytes].clone()
                else:
                    assert torch.equal(ref, packed[:nbytes]), "kernels disagree"
                ms_d, z_hat = timed(lambda: eb.decompress_from_device(packed, offs, lens, bias=comp.biasing, exp_scale=es), iters)
                z_hat = z_hat[0] if isinstance(z_hat, tuple) else z_hat
                if zref is None:
                    zref = z_hat.clone()
                else:
                    assert torch.equal(zref, z_hat), "decoders disagree"
                a_e = rows * 512 * 2 + nbytes + 4 * rows
                a_d = nbytes + rows * 512 * 4
                lines.append(f"{tag:10s} {rows:7d} {name:>6s} {ms_e * 1e3:12.1f} {a_e / ms_e / 1e6:8.1f} {100 * a_e / ms_e / 1e6 / hbm:6.2f} "
                             f"{ms_d * 1e3:10.1f} {a_d / ms_d / 1e6:8.1f} {100 * a_d / ms_d / 1e6 / hbm:6.2f} {8.0 * nbytes / rows:9.1f}")
                print(lines[-1], flush=True)
            del z, zref
    N.call("ll_rans_set_impl", 0)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
