"""Stream-per-lane encode + decode at 131 072 rows, twice, for ncu captures:
    ncu --set full -k regex:'lanes' -c 2 ... python tools/ncu_codec.py"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf  # noqa: E402
from lossyless_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")
comp, _ = hubconf.clip_compressor_b005(device=dev)
eb = comp.entropy_bottleneck
g = torch.Generator(device=dev).manual_seed(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
z16 = (torch.randn(rows, 512, generator=g, device=dev) * 0.45).half()
for _ in range(2):
    N.call("ll_rans_set_impl", 2)
    packed, offsets, lens, _ = eb.compress_to_device(z16, bias=comp.biasing, exp_scale=comp._exp_scale(), mode=0)
    eb.decompress_from_device(packed, offsets[:-1].contiguous(), lens, bias=comp.biasing, exp_scale=comp._exp_scale())
N.call("ll_rans_set_impl", 0)
torch.cuda.synchronize()
