"""The memory-bound kernels of a block's backward pass alone, at batch 256 (12 800 token rows), bf16, inputs rotated
through more buffers than the 126 MB L2 holds ("cold") or not ("warm"): ll_colsum (bias gradients) on its four shapes,
ll_layernorm_backward (3 reads + 1 write of [rows, 768]).
    python tools/bwd_kernels_bench.py"""
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

print("# ll_layernorm_backward, bf16 [rows, 768]: x, g_h, g_res read, g_x written, g_gamma / g_beta accumulated")
for R in (12800, 256):
    W = 768
    nbuf = max(2, int(400e6 // (R * W * 2 * 4)) + 1)
    bufs = [[torch.randn(R, W, device=dev, dtype=torch.bfloat16) for _ in range(4)] for _ in range(nbuf)]
    gamma = torch.rand(W, device=dev) + 0.5
    gw, gb = torch.zeros(W, device=dev), torch.zeros(W, device=dev)

    def run(b):
        x, gh, gr, gx = b
        N.call("ll_layernorm_backward", N.ptr(x), N.ptr(gh), N.ptr(gamma), N.ptr(gr), N.ptr(gx), N.ptr(gw), N.ptr(gb), R, W,
               N.dtype_code(x), N.stream())

    # reference: autograd through torch's float32 layer_norm
    x, gh, gr, gx = bufs[0]
    run(bufs[0])
    xf = x.float().requires_grad_(True)
    gam = gamma.clone().requires_grad_(True)
    beta = torch.zeros(W, device=dev, requires_grad=True)
    torch.nn.functional.layer_norm(xf, (W,), gam, beta, 1e-5).backward(gh.float())
    e_x = (gx.float() - (xf.grad + gr.float())).abs().max().item()
    e_w = (gw - gam.grad).abs().max().item() / gam.grad.abs().max().item()
    e_b = (gb - beta.grad).abs().max().item() / beta.grad.abs().max().item()
    for warm in (False, True):
        for _ in range(3):
            run(bufs[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10 * nbuf
        e0.record()
        for i in range(iters):
            run(bufs[0] if warm else bufs[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print(f"R={R:6d} {'warm' if warm else 'cold'}: {us:7.2f} us  {R * W * 2 * 4 / us / 1e6:7.2f} TB/s  "
              f"max |d g_x| {e_x:.1e} (bf16 output)  rel err g_gamma {e_w:.1e} g_beta {e_b:.1e}")
