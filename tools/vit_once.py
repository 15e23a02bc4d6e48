"""Two ViT-B/32 forwards of one encoder pass (the second one is what ncu captures).
usage: python tools/vit_once.py [images]   (default 1512 = the library's pass size)"""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1512
comp, _ = hubconf.clip_compressor_b005(device="cuda")
x = torch.randn(n, 3, 224, 224, device="cuda").half()
for _ in range(2):
    comp.compress(x)
torch.cuda.synchronize()
