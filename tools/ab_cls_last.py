"""A/B on one box: the last block's row-wise half on the class-token rows only (default) vs on every row."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf  # noqa: E402

comp, _ = hubconf.clip_compressor_b005(device="cuda")
vit = comp.clip
for n in (32, 64, 128, 256, 512, 1512, 3024):
    x = torch.randn(n, 3, 224, 224, device="cuda").half()
    row = f"batch {n:5d}:"
    for on in (1, 0, 1, 0):
        vit.set_cls_last_block(on)
        for _ in range(3):
            vit(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if n <= 512 else 8
        e0.record()
        for _ in range(reps):
            vit(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row += f"  {'cls' if on else 'all'} {ms:8.3f} ms {n / ms:8.1f} k img/s"
    print(row)
