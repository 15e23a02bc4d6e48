"""One launch each of the entropy kernels at the sizes the design notes quote, for ncu captures:
EB likelihood forward / backward (131 072 rows), stream-per-lane encode / decode (131 072 rows), stream-per-warp encode
(1512 rows, the bench step's codec tail), Gaussian-conditional likelihood forward (131 072 rows).
    ncu --set full -k regex:'eb_forward|eb_backward|lanes|eb_rans_encode_kernel|gc_forward' ... python tools/ncu_entropy.py"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf  # noqa: E402
from lossyless_b200 import entropy_models as em  # noqa: E402

dev = torch.device("cuda:0")
comp, _ = hubconf.clip_compressor_b005(device=dev)
eb = comp.entropy_bottleneck
g = torch.Generator(device=dev).manual_seed(0)
rows = 131072
z = (torch.randn(rows, 512, generator=g, device=dev) * 0.45)
z16 = z.half()
for _ in range(2):
    with torch.no_grad():
        em.eb_rate_forward(eb, z, None, comp.biasing, comp.scaling, want_rates=True)
    packed, offsets, lens, _ = eb.compress_to_device(z16, bias=comp.biasing, exp_scale=comp._exp_scale(), mode=0)
    eb.decompress_from_device(packed, offsets[:-1].contiguous(), lens, bias=comp.biasing, exp_scale=comp._exp_scale())
    eb.compress_to_device(z16[:1512], bias=comp.biasing, exp_scale=comp._exp_scale(), mode=1)
    gc = em.GaussianConditional(None).to(dev).train()
    with torch.no_grad():
        gc.forward_rates(z, torch.rand_like(z) + 0.2, means=torch.zeros_like(z))
    # training mode forward + backward of the bottleneck (2048 rows: the training size; and the large size)
    for n in (2048, rows):
        zz = z[:n].clone().requires_grad_(True)
        noise = torch.rand(n, 512, device=dev) - 0.5
        _, _, r = em.eb_rate_forward(eb, zz, noise, comp.biasing, comp.scaling)
        r.sum().backward()
torch.cuda.synchronize()
