"""Per-shape timing of the tcgen05 GEMM at the encoder's shapes (CUDA events, L2-sized rotation).
usage: python tools/gemm_bench.py [images_per_pass]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lossyless_b200 import _native as N  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
shapes = [("patch", n_img * 49, 768, 3072, 0, False), ("qkv", n_img * 50, 2304, 768, 0, False),
          ("out", n_img * 50, 768, 768, 0, True), ("fc1", n_img * 50, 3072, 768, 1, False),
          ("fc2", n_img * 50, 768, 3072, 0, True), ("proj", n_img, 512, 768, 0, False)]
tot_ms, tot_fl = 0.0, 0.0
for name, M, Nn, K, act, res in shapes:
    nbuf = 3  # rotate buffers so A/D are not trivially L2 resident from the previous launch
    As = [torch.randn(M, K, device=dev).half() for _ in range(nbuf)]
    W = (torch.randn(Nn, K, device=dev) * K ** -0.5).half()
    bias = torch.randn(Nn, device=dev)
    Ds = [torch.zeros(M, Nn, device=dev).half() for _ in range(nbuf)]

    def run(i):
        A, D = As[i % nbuf], Ds[i % nbuf]
        N.call("ll_gemm", N.ptr(A), N.ptr(W), N.ptr(bias), N.ptr(D) if res else None, N.ptr(D), M, Nn, K, 1, act, N.stream())

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * M * Nn * K
    mult = 12 if name in ("qkv", "out", "fc1", "fc2") else 1
    tot_ms += ms * mult
    tot_fl += fl * mult
    print(f"{name:6s} M={M:6d} N={Nn:5d} K={K:5d}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  (x{mult})")
print(f"per pass of {n_img} images: GEMM {tot_ms:.3f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s, {n_img / tot_ms * 1e3:.0f} img/s GEMM-only")
