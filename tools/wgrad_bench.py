"""Weight-gradient GEMMs of one ViT block (D = dY^T X, reduction over the tokens) at the training step's size:
ll_gemm_tn (single pass) against ll_gemm_tn_splitk (float32 partials + reduce), CUDA events, rotating operands.
usage: python tools/wgrad_bench.py [tokens]   (default 25 600 = 2 views x 256 images x 50 tokens)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lossyless_b200 import _native as N  # noqa: E402

Kt = int(sys.argv[1]) if len(sys.argv) > 1 else 25600
dev = torch.device("cuda")
W = 768
shapes = [("c_proj", W, 4 * W), ("c_fc", 4 * W, W), ("in_proj", 3 * W, W), ("out_proj", W, W)]
ws_bytes = 64 * W * W  # what ll_vit_block_backward lends
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)


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


print(f"# weight-gradient GEMMs, {Kt} tokens, bf16, workspace {ws_bytes / 1e6:.1f} MB")
tot = {"single": 0.0, "splitk": 0.0}
for name, Mo, No in shapes:
    nbuf = 3
    As = [torch.randn(Kt, Mo, device=dev).bfloat16() for _ in range(nbuf)]
    Bs = [torch.randn(Kt, No, device=dev).bfloat16() for _ in range(nbuf)]
    D = torch.empty(Mo, No, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * Kt * Mo * No
    row = f"{name:9s} [{Mo:4d} x {No:4d}]"
    for label, pair in (("single", -1), ("splitk", -1), ("splitk/no-pair", 0)):
        N.call("ll_gemm_set_cta_pair", pair)
        if label == "single":
            us = timeit(lambda i: N.call("ll_gemm_tn", N.ptr(As[i % nbuf]), N.ptr(Bs[i % nbuf]), N.ptr(D), Kt, Mo, No, 2, N.stream()))
        else:
            us = timeit(lambda i: N.call("ll_gemm_tn_splitk", N.ptr(As[i % nbuf]), N.ptr(Bs[i % nbuf]), N.ptr(D), Kt, Mo, No, 2,
                                         N.ptr(ws), ws_bytes, N.stream()))
        if label in tot:
            tot[label] += us
        row += f"   {label} {us:7.1f} us {fl / us / 1e6:7.1f} TFLOP/s"
    N.call("ll_gemm_set_cta_pair", -1)
    print(row)
print(f"per block: single pass {tot['single']:.1f} us, split-K {tot['splitk']:.1f} us")
