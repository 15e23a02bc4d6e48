"""SURVEY.md 8(d) configuration 5: the encode sweep  beta in {0.01, 0.05, 0.1} x batch in {32 .. 2048}
on one GPU, inputs N(0,1) fp16 [B,3,224,224] resident on the device, >= 3 warm-ups, >= 20 timed
iterations, CUDA events.  Also the bottleneck stage in isolation (z ~ N(0, 0.45^2) fed to EB quantise +
rANS + pack), whose bits/img has a known expectation (~1603 / 2055 / 3557 + 32 header bits), and the
large-row HBM figure of the codec and likelihood kernels (rows >= 65 536).

usage: python tools/sweep.py [--iters 20] [--out profiles/rNN_sweep.txt]
"""
import argparse
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf  # noqa: E402
from lossyless_b200 import entropy_models as em  # noqa: E402

ENTRIES = [("beta1e-02", hubconf.clip_compressor_b001), ("beta5e-02", hubconf.clip_compressor_b005), ("beta1e-01", hubconf.clip_compressor_b01)]
BATCHES = [32, 64, 128, 256, 512, 1024, 2048]


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lines = [f"# encode sweep on {torch.cuda.get_device_name(0)}; inputs device-resident; {args.iters} timed iterations after 3 warm-ups",
             "# compress = ViT-B/32 (seeded random weights) + EB quantise + rANS + pack, records left on the device",
             f"{'ckpt':10s} {'batch':>6s} {'ms/batch':>9s} {'img/s':>10s} {'bits/img':>9s}"]
    g = torch.Generator(device=dev).manual_seed(0)
    x_all = torch.randn(2048, 3, 224, 224, generator=g, device=dev, dtype=torch.float16)
    for tag, entry in ENTRIES:
        comp, _ = entry(device=dev)
        for B in BATCHES:
            x = x_all[:B]
            ms, (blob, offs, lens) = timed(lambda: comp._compress_records(x, 1), args.iters)
            lines.append(f"{tag:10s} {B:6d} {ms:9.3f} {B / ms * 1e3:10.0f} {8.0 * float(lens.sum()) / B:9.1f}")
        # bottleneck stage alone at a known distribution
        z = (torch.randn(2048, 512, generator=g, device=dev) * 0.45).half()
        enc = lambda zz: comp.entropy_bottleneck.compress_to_device(zz, bias=comp.biasing, exp_scale=comp._exp_scale(), mode=1)  # noqa: E731
        ms, (blob, offs, lens, _) = timed(lambda: enc(z), args.iters)
        lines.append(f"{tag:10s} EB+rANS only, z ~ N(0, 0.45^2), 2048 rows: {ms * 1e3:8.1f} us  {8.0 * float(lens.sum()) / 2048 + 32:8.1f} bits/img incl. 32-bit record header")
        print("\n".join(lines[-len(BATCHES) - 1:]), flush=True)
    # large rows: the HBM figure of the bottleneck kernels (SURVEY 8d "honest caveat")
    comp, _ = hubconf.clip_compressor_b005(device=dev)
    rows = 131072
    z = (torch.randn(rows, 512, generator=g, device=dev) * 0.45).half()
    ms, (blob, offs, lens, _) = timed(lambda: comp.entropy_bottleneck.compress_to_device(z, bias=comp.biasing, exp_scale=comp._exp_scale(), mode=1), 10)
    algo = rows * 512 * 2 + int(lens.sum()) + 4 * rows
    lines.append(f"large rows: EB quantise + rANS + pack, {rows} rows: {ms:.3f} ms, algorithmic {algo / 1e6:.1f} MB -> {algo / ms / 1e6:.1f} GB/s")
    eb = comp.entropy_bottleneck
    zf = z.float()
    with torch.no_grad():
        ms, _ = timed(lambda: em.eb_rate_forward(eb, zf, None, comp.biasing, comp.scaling, want_rates=True), 10)
    algo = rows * 512 * 8 + rows * 4
    lines.append(f"large rows: EB likelihood forward (z_hat + per-row rate), {rows} rows: {ms:.3f} ms, algorithmic {algo / 1e6:.1f} MB -> {algo / ms / 1e6:.1f} GB/s")
    print("\n".join(lines[-2:]))
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
