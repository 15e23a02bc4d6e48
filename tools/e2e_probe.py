"""Where the end-to-end encode time goes: preprocessing kernels alone, the device-resident step alone, the same
step with an unrelated H2D stream running beside it (link / L2 / power contention), and the full pipelines.
usage: python tools/e2e_probe.py [batch]"""
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf  # noqa: E402
from lossyless_b200 import _native as N  # noqa: E402
from lossyless_b200.preprocess import GpuClipPreprocess  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1512
dev = torch.device("cuda:0")
comp, _ = hubconf.clip_compressor_b005(device=dev)


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


raw = torch.randint(0, 256, (B, 96, 96, 3), dtype=torch.uint8, device=dev)
pre = GpuClipPreprocess(224)
for impl in (0, 1):
    N.call("ll_clip_preprocess_set_impl", impl)
    print(f"preprocess {B} x 96x96 -> 224x224 fp16, impl {impl}: {timed(lambda: pre(raw)):.3f} ms")
N.call("ll_clip_preprocess_set_impl", 0)

X = torch.randn(B, 3, 224, 224, device=dev).half()
step = lambda: comp._compress_records(X, 1)  # noqa: E731
print(f"device-resident step: {timed(step):.3f} ms")

X_host = X.cpu().pin_memory()
side = torch.cuda.Stream(dev)
dummy = torch.empty_like(X)
stop = [False]


def step_with_copy():
    with torch.cuda.stream(side):
        dummy.copy_(X_host, non_blocking=True)
    step()


print(f"device-resident step + concurrent 455 MB H2D on a side stream: {timed(step_with_copy):.3f} ms")
side.synchronize()
raw_dev_step = lambda: comp._compress_records(pre(raw), 1)  # noqa: E731
print(f"device-resident raw uint8 -> preprocess -> step: {timed(raw_dev_step):.3f} ms")


def wall(fn, iters=10):
    fn(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(iters)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def stream_run(x):
    def run(k):
        for _ in comp.compress_stream(x for _ in range(k)):
            pass
    return run


raw_host = raw.cpu().pin_memory()
print(f"compress_stream, pinned fp16 host batches: {wall(stream_run(X_host)):.3f} ms/batch")
print(f"compress_stream, pinned raw uint8 host batches: {wall(stream_run(raw_host)):.3f} ms/batch")
for sb in (189, 756, 1512):
    def run(k, sb=sb):
        for _ in comp.compress_stream((X_host for _ in range(k)), sub_batch=sb):
            pass
    print(f"compress_stream fp16, sub_batch {sb}: {wall(run):.3f} ms/batch")

# how long the 455 MB H2D copy itself takes while the encoder runs (copy timed with events on its own stream)
def copy_ms_under_load():
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        c0.record(side)
        dummy.copy_(X_host, non_blocking=True)
        c1.record(side)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    return c0.elapsed_time(c1)


t = sorted(copy_ms_under_load() for _ in range(5))[2]
print(f"455 MB pinned H2D while the encoder runs: {t:.3f} ms -> {X_host.numel() * 2 / t / 1e6:.1f} GB/s")
torch.cuda.synchronize()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record()
dummy.copy_(X_host, non_blocking=True)
c1.record()
torch.cuda.synchronize()
print(f"455 MB pinned H2D alone: {c0.elapsed_time(c1):.3f} ms -> {X_host.numel() * 2 / c0.elapsed_time(c1) / 1e6:.1f} GB/s")
for sb in (378, 504):
    def run(k, sb=sb):
        for _ in comp.compress_stream((X_host for _ in range(k)), sub_batch=sb):
            pass
    print(f"compress_stream fp16, sub_batch {sb}: {wall(run):.3f} ms/batch")
