"""One ViT-B/32 encoder pass (1512 images by default) for ncu captures:  ncu ... python tools/ncu_one_pass.py [batch] [passes]
Launch order per pass: patch-embed GEMM, assemble+ln_pre, then per block qkv GEMM, attention, out-proj GEMM, c_fc GEMM,
c_proj GEMM; ln_post, proj GEMM; then the bottleneck + rANS + pack kernels of compress()."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import bench  # noqa: E402
import hubconf  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1512
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
comp, _ = hubconf.clip_compressor_b005(device="cuda:0")
X = bench.bench_images(B).to("cuda:0")
for _ in range(passes):
    comp._compress_records(X, mode=1)
torch.cuda.synchronize()
