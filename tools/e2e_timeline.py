"""Timeline of compress_stream's pipeline: CUDA-event timestamps of every batch's H2D copy (side stream) and
encode (main stream), to see what overlaps.  usage: python tools/e2e_timeline.py [batch] [sub_batch] [u8]"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1512
SB = int(sys.argv[2]) if len(sys.argv) > 2 else 1512
dev = torch.device("cuda:0")
comp, _ = hubconf.clip_compressor_b005(device=dev)
U8 = len(sys.argv) > 3 and sys.argv[3] == "u8"  # the CLIP-sized uint8 feed of bench.py's e2e instead of fp16 tensors
X_host = (torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8) if U8 else torch.randn(B, 3, 224, 224).half()).pin_memory()
for _ in comp.compress_stream((X_host for _ in range(3)), sub_batch=SB):
    pass
torch.cuda.synchronize()

marks = []
orig = comp._encode_host_batch


def traced(X, sub_batch, *args, **kw):
    pipe = comp._h2d_pipe(dev, X, sub_batch)
    main = torch.cuda.current_stream(dev)
    a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    a.record(pipe["stream"])      # side stream reaches this batch
    c.record(main)                # main stream reaches this batch
    out = orig(X, sub_batch, *args, **kw)
    b.record(pipe["stream"])      # last H2D of the batch done
    d.record(main)                # encode of the batch done
    marks.append((a, b, c, d))
    return out


comp._encode_host_batch = traced
t0 = torch.cuda.Event(enable_timing=True)
t0.record()
import time

host = []
th = time.perf_counter()
for _ in comp.compress_stream((X_host for _ in range(8)), sub_batch=SB):
    host.append(time.perf_counter() - th)
torch.cuda.synchronize()
print("host: batch i handed back at ms", [round(1e3 * h, 2) for h in host])
print(f"batch {B}, sub_batch {SB}: times in ms since start")
for i, (a, b, c, d) in enumerate(marks):
    print(f"  batch {i}: H2D stream {t0.elapsed_time(a):7.2f} -> {t0.elapsed_time(b):7.2f}   main stream {t0.elapsed_time(c):7.2f} -> {t0.elapsed_time(d):7.2f}")
