"""Every (CTA pair?, tile width) candidate of the tcgen05 GEMM at the encoder's shapes and a range of row counts, timed
with CUDA events, next to what pick_tile's time model chooses by itself.  usage: python tools/tile_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lossyless_b200 import _native as N  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


shapes = [("qkv", 2304, 768, 0, False), ("out", 768, 768, 0, True), ("fc1", 3072, 768, 1, False), ("fc2", 768, 3072, 0, True)]
for n_img in (32, 64, 128, 256, 512, 1512):
    M = n_img * 50
    for name, Nn, K, act, res in shapes:
        nbuf = 3
        As = [torch.randn(M, K, device=dev).half() for _ in range(nbuf)]
        W = (torch.randn(Nn, K, device=dev) * K ** -0.5).half()
        bias = torch.randn(Nn, device=dev)
        Ds = [torch.zeros(M, Nn, device=dev).half() for _ in range(nbuf)]

        def run(i):
            A, D = As[i % nbuf], Ds[i % nbuf]
            N.call("ll_gemm", N.ptr(A), N.ptr(W), N.ptr(bias), N.ptr(D) if res else None, N.ptr(D), M, Nn, K, 1, act, N.stream())

        N.call("ll_gemm_set_cta_pair", -1)
        N.call("ll_gemm_set_tile_n", 0)
        auto = timeit(run)
        row, best = "", (1e9, None)
        for pair in (0, 1):
            for bn in (256, 192, 128, 64):
                if Nn % bn or (pair and (bn == 64 or M <= 128)):
                    continue
                N.call("ll_gemm_set_cta_pair", pair)
                N.call("ll_gemm_set_tile_n", bn)
                us = timeit(run)
                row += f" {'P' if pair else 'S'}{bn}:{us:6.1f}"
                if us < best[0]:
                    best = (us, f"{'P' if pair else 'S'}{bn}")
        N.call("ll_gemm_set_cta_pair", -1)
        N.call("ll_gemm_set_tile_n", 0)
        print(f"{n_img:5d} img {name:4s} M={M:6d}: auto {auto:6.1f} us | best {best[1]:>4s} {best[0]:6.1f} us ({auto / best[0]:.2f}x) |{row}")
