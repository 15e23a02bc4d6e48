"""Where a training step's time goes: GPU kernel time (CUPTI timestamps through torch.profiler) against the step's span on the
device, the largest idle gaps between consecutive kernels, and the host-side enqueue time per step.

    python tools/train_timeline.py [--batch 256] [--steps 3]
"""
import argparse
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--splitk", type=int, default=1, help="0: weight gradients in one pass (A/B of ll_gemm_tn_splitk)")
    ap.add_argument("--brief", action="store_true", help="skip the host-side profiler table")
    ap.add_argument("--fusion", type=lambda v: int(v, 0), default=0x7FFFFFFF, help="ll_vit_block_set_fusion mask (A/B)")
    args = ap.parse_args()
    from lossyless_b200 import _native

    _native.call("ll_gemm_set_splitk", args.splitk)
    _native.load().ll_vit_block_set_fusion(args.fusion)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    hp = default_clip_hparams(distortion="BINCE", beta=1e-2, max_steps=1000, n_data=50000, is_frozen=False)
    model = LearnableCompressor(hp).to(dev).train()
    B = args.batch
    batch = (torch.randn(B, 3, 224, 224, device=dev), (None, torch.randn(B, 3, 224, 224, device=dev)))
    model.fit([batch] * 3)
    torch.cuda.synchronize()
    # host enqueue time vs device time of one step (no profiler)
    opts = model._trainer_state["optimizers"]
    model._defer_log_sync = True
    e = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    host = []
    e[0].record()
    for i in range(args.steps):
        t0 = time.perf_counter()
        model._fit_loop([batch], opts, None, 3.0, False, None, None)
        host.append(time.perf_counter() - t0)
        e[i + 1].record()
    torch.cuda.synchronize()
    print("host enqueue ms per step:", [round(1e3 * h, 2) for h in host])
    print("device ms per step      :", [round(e[i].elapsed_time(e[i + 1]), 2) for i in range(args.steps)])
    # every host synchronisation of one step, with the Python line that caused it
    import traceback
    import collections

    sites = collections.Counter()
    orig = warnings.showwarning

    def show(message, category, filename, lineno, file=None, line=None):
        if "synchroniz" in str(message):
            st = [f for f in traceback.extract_stack()[:-1] if "/torch/" not in f.filename]
            sites[" <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in reversed(st[-6:]))] += 1

    warnings.showwarning = show
    warnings.simplefilter("always")
    torch.cuda.set_sync_debug_mode(1)
    model._fit_loop([batch], opts, None, 3.0, False, None, None)
    torch.cuda.set_sync_debug_mode(0)
    warnings.showwarning = orig
    warnings.filterwarnings("ignore")
    print("synchronising calls in one step:", sum(sites.values()))
    for k, v in sites.most_common(30):
        print(f"  {v:3d} x {k}")
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model._fit_loop([batch] * args.steps, opts, None, 3.0, False, None, None)
        torch.cuda.synchronize()
    ev = [x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA and x.time_range is not None]
    ev.sort(key=lambda x: x.time_range.start)
    t0, t1 = ev[0].time_range.start, max(x.time_range.end for x in ev)
    busy = sum(x.time_range.end - x.time_range.start for x in ev)
    print(f"{len(ev)} device activities over {args.steps} steps: span {(t1 - t0) / 1e3 / args.steps:.2f} ms/step, busy {busy / 1e3 / args.steps:.2f} ms/step")
    gaps = []
    end = ev[0].time_range.end
    for a in ev[1:]:
        if a.time_range.start > end:
            gaps.append((a.time_range.start - end, a.name[:60]))
        end = max(end, a.time_range.end)
    gaps.sort(reverse=True)
    print("idle total ms/step:", round(sum(g for g, _ in gaps) / 1e3 / args.steps, 2), " gaps > 20 us:", sum(1 for g, _ in gaps if g > 20) // args.steps, "per step")
    for g, n in gaps[:25]:
        print(f"  {g:8.1f} us before {n}")
    # what the host was doing during the three largest gaps (runtime API calls and operators overlapping the window)
    cpu = [x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CPU and x.time_range is not None]
    big = []
    end = ev[0].time_range.end
    for a in ev[1:]:
        if a.time_range.start > end:
            big.append((a.time_range.start - end, end, a.time_range.start))
        end = max(end, a.time_range.end)
    big.sort(reverse=True)
    for g, t_a, t_b in big[:3]:
        print(f"  host activity during the {g:.0f} us gap:")
        hits = [(min(c.time_range.end, t_b) - max(c.time_range.start, t_a), c) for c in cpu if c.time_range.start < t_b and c.time_range.end > t_a]
        hits.sort(key=lambda h: -h[0])
        for ov, c in hits[:12]:
            print(f"      {ov:8.1f} us of {c.name[:70]} (total {c.time_range.end - c.time_range.start:.0f} us)")
    # every synchronising runtime call inside the profiled steps, with the operators enclosing it
    for c in cpu:
        if "Synchronize" in c.name or c.name in ("cudaFree", "cudaMalloc", "cudaHostAlloc", "cudaMemcpy"):
            enc = [o for o in cpu if o is not c and o.time_range.start <= c.time_range.start and o.time_range.end >= c.time_range.end]
            enc.sort(key=lambda o: o.time_range.end - o.time_range.start)
            print(f"  {c.name} at +{(c.time_range.start - t0) / 1e3:.2f} ms for {c.time_range.end - c.time_range.start:.0f} us inside: "
                  + " < ".join(o.name[:40] for o in enc[:4]))
    agg = {}
    for x in ev:
        k = x.name.split("(")[0].replace("void ", "")[:60]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += x.time_range.end - x.time_range.start
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"{t / 1e3 / args.steps:9.3f} ms {c // args.steps:5d}  {k}")
    if not args.brief:
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=35, max_name_column_width=60))


if __name__ == "__main__":
    main()
