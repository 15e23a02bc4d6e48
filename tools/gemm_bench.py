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
DBG = int(os.environ.get("LL_DBG", "0"), 0)
pair_mode = int(os.environ.get("LL_PAIR", "-1"))
N.call("ll_gemm_set_cta_pair", pair_mode)
N.call("ll_gemm_set_tile_n", int(os.environ.get("LL_BN", "0")))
print(f"--- cta_pair mode {pair_mode}, {n_img} images per pass, dbg {DBG:#x}")
tot_ms, tot_fl = 0.0, 0.0
for name, M, Nn, K, act, res in shapes:
    nbuf = 3  # rotate buffers so A/D are not trivially L2 resident from the previous launch
    As = [torch.randn(M, K, device=dev).half() for _ in range(nbuf)]
    W = (torch.randn(Nn, K, device=dev) * K ** -0.5).half()
    bias = torch.randn(Nn, device=dev)
    Ds = [torch.zeros(M, Nn, device=dev).half() for _ in range(nbuf)]

    def run(i):
        A, D = As[i % nbuf], Ds[i % nbuf]
        N.call("ll_gemm", N.ptr(A), N.ptr(W), N.ptr(bias), N.ptr(D) if res else None, N.ptr(D), M, Nn, K, 1, act | DBG, N.stream())

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
    extra = ""
    if os.environ.get("LL_CUBLAS"):
        # library yard-stick on the same shape: cuBLASLt through torch (bias epilogue only; the activation /
        # residual / LayerNorm-statistics work our epilogue fuses would be extra launches on top of this)
        def ref(i):
            torch.nn.functional.linear(As[i % nbuf], W, bias.half())

        for i in range(3):
            ref(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            ref(i)
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / reps
        cub_ms = globals().get("cub_ms", 0.0) + ms_ref * mult
        globals()["cub_ms"] = cub_ms
        extra = f"   cuBLAS {ms_ref * 1e3:8.1f} us {fl / ms_ref / 1e9:7.1f} TFLOP/s"
    print(f"{name:6s} M={M:6d} N={Nn:5d} K={K:5d}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  (x{mult}){extra}")
print(f"per pass of {n_img} images: GEMM {tot_ms:.3f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s, {n_img / tot_ms * 1e3:.0f} img/s GEMM-only")
if os.environ.get("LL_CUBLAS"):
    print(f"cuBLAS (bias epilogue only) on the same shapes: {cub_ms:.3f} ms, {tot_fl / cub_ms / 1e9:.1f} TFLOP/s")

if os.environ.get("LL_GEMM_ONLY"):
    sys.exit(0)

# ---- the encoder's non-GEMM kernels and the codec at the same pass size
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


M = n_img * 50
qkv = torch.randn(M, 2304, device=dev).half()
o = torch.empty(M, 768, device=dev).half()
us = timeit(lambda: N.call("ll_attention", N.ptr(qkv), N.ptr(o), n_img, 50, 12, 1, N.stream()))
print(f"attention  B={n_img}: {us:8.1f} us  ({(qkv.numel() + o.numel()) * 2 / us / 1e3:.0f} GB/s)  (x12)")
x = torch.randn(M, 768, device=dev).half()
h = torch.empty_like(x)
g = torch.ones(768, device=dev)
b = torch.zeros(768, device=dev)
us_ln = timeit(lambda: N.call("ll_layernorm", N.ptr(x), M, 768, N.ptr(g), N.ptr(b), N.ptr(h), 1, N.stream()))
print(f"layernorm  rows={M}: {us_ln:8.1f} us  ({x.numel() * 4 / us_ln / 1e3:.0f} GB/s)  (x24)")

import warnings
warnings.filterwarnings("ignore")
import hubconf
comp, _ = hubconf.clip_compressor_b005(device="cuda")
for rows in (256, 1024, 4096):
    z = (torch.randn(rows, 512, device=dev) * 1.0).half()
    us_e = timeit(lambda: comp.entropy_bottleneck.compress_to_device(z, comp.biasing, comp._exp_scale(), mode=1), reps=10)
    zs = (torch.randn(rows, 512, device=dev) * 0.45).half()
    us_s = timeit(lambda: comp.entropy_bottleneck.compress_to_device(zs, comp.biasing, comp._exp_scale(), mode=1), reps=10)
    print(f"eb+rans encode+pack rows={rows}: {us_e:8.1f} us (z std 1.0)   {us_s:8.1f} us (z std 0.45)")
X = torch.randn(n_img, 3, 224, 224, device=dev).half()
us_v = timeit(lambda: comp.clip(X), reps=10)
print(f"vit forward B={n_img}: {us_v:8.1f} us -> {n_img / us_v * 1e6:.0f} img/s")
