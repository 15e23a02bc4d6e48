"""The shared-memory kernels added / changed at the end of round 2, launched alone (no torch reference kernels in the
process) for compute-sanitizer --tool racecheck: ll_layernorm_backward, ll_colsum, ll_embed_tokens, ll_embed_backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lossyless_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for dt in (torch.bfloat16, torch.float16):
    R, W = 2 * 50 * 5, 768
    x, gh, gr = (torch.randn(R, W, generator=g).to(dev, dt) for _ in range(3))
    gx = torch.empty_like(x)
    gamma, gw, gb = torch.rand(W, generator=g).to(dev) + 0.5, torch.zeros(W, device=dev), torch.zeros(W, device=dev)
    N.call("ll_layernorm_backward", N.ptr(x), N.ptr(gh), N.ptr(gamma), N.ptr(gr), N.ptr(gx), N.ptr(gw), N.ptr(gb), R, W, N.dtype_code(x), N.stream())
    for C in (768, 2304, 3072):
        t = torch.randn(777, C, generator=g).to(dev, dt)
        out = torch.zeros(C, device=dev)
        N.call("ll_colsum", N.ptr(t), 777, C, N.dtype_code(t), N.ptr(out), N.stream())
    B, S = 37, 50
    patch, cls, pos = torch.randn(B, S - 1, W, generator=g).to(dev, dt), torch.randn(W, generator=g).to(dev, dt), torch.randn(S, W, generator=g).to(dev, dt)
    tok = torch.empty(B, S, W, device=dev, dtype=dt)
    N.call("ll_embed_tokens", N.ptr(patch), N.ptr(cls), N.ptr(pos), N.ptr(tok), B, S, W, N.dtype_code(tok), N.stream())
    g_patch, g_pos, g_cls = torch.empty_like(patch), torch.zeros(S, W, device=dev), torch.zeros(W, device=dev)
    N.call("ll_embed_backward", N.ptr(tok), N.ptr(g_patch), N.ptr(g_pos), N.ptr(g_cls), B, S, W, N.dtype_code(tok), N.stream())
torch.cuda.synchronize()
print("launched")
