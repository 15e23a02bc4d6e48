"""Column sums (bias gradients of the training step) alone: ll_colsum on the four shapes of a block's backward pass
at batch 256 (12 800 token rows), bf16, inputs rotated through more buffers than the 126 MB L2 holds.
    python tools/colsum_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lossyless_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")
print(f"# ll_colsum on {torch.cuda.get_device_name(0)}; bf16 [12800, C]; cold = inputs rotated through > 126 MB")
for R, C in ((12800, 768), (12800, 2304), (12800, 3072), (256, 768), (256, 3072)):
    nbuf = max(2, int(400e6 // (R * C * 2)) + 1)
    xs = [torch.randn(R, C, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    out = torch.zeros(C, device=dev, dtype=torch.float32)
    N.call("ll_colsum", N.ptr(xs[0]), R, C, N.dtype_code(xs[0]), N.ptr(out), N.stream())
    ref = xs[0].float().sum(0)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    for warm in (False, True):
        for _ in range(3):
            N.call("ll_colsum", N.ptr(xs[0]), R, C, N.dtype_code(xs[0]), N.ptr(out), N.stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10 * nbuf
        e0.record()
        for i in range(iters):
            x = xs[0] if warm else xs[i % nbuf]
            N.call("ll_colsum", N.ptr(x), R, C, N.dtype_code(x), N.ptr(out), N.stream())
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print(f"R={R:6d} C={C:5d} {'warm' if warm else 'cold'}: {us:7.2f} us  {R * C * 2 / us / 1e6:7.2f} TB/s  rel err {err:.1e}")
