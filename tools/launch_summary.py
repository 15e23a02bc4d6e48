"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> per-kernel totals of ONE step.
    python tools/launch_summary.py launches.csv [--first REGEX] [--occurrence K]
The step window runs from the K-th launch matching --first (default: the second occurrence of the first kernel of the
list's steady state) to the next one.  Times under ncu are cold-cache and serialised: compare SHARES."""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--first", default="adamw_flat_kernel", help="regex of the kernel that ends a step")
    ap.add_argument("--occurrence", type=int, default=-2, help="which match closes the window (python index into the matches)")
    ap.add_argument("--stride", type=int, default=1, help="matches per step (e.g. 2 adamw launches per training step)")
    a = ap.parse_args()
    with open(a.csv) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(x["Kernel Name"], float(x["Metric Value"]), x["Metric Unit"]) for x in csv.DictReader(lines) if x.get("Metric Name") == "gpu__time_duration.sum"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}[rows[0][2]]
    idx = [i for i, r in enumerate(rows) if re.search(a.first, r[0])]
    ends = idx[a.stride - 1::a.stride]
    hi = ends[a.occurrence] + 1
    lo = ends[a.occurrence - 1] + 1
    win = rows[lo:hi]
    agg = collections.OrderedDict()
    for n, v, _ in win:
        n = re.sub(r"^void ", "", n.split("(")[0])[:72]
        e = agg.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += v * scale
    tot = sum(e[1] for e in agg.values())
    print(f"# launches {lo}..{hi - 1} of {len(rows)}: {len(win)} launches, {tot / 1e3:.2f} ms summed (cold-cache, serialised)")
    print(f"{'us':>10s} {'launches':>8s} {'share':>6s}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot >= 0.0005:
            print(f"{t:10.1f} {c:8d} {100 * t / tot:5.1f}%  {n}")


if __name__ == "__main__":
    main()
