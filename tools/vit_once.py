"""Two ViT-B/32 forwards of one 378-image pass (the second one is what ncu captures)."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import hubconf
comp, _ = hubconf.clip_compressor_b005(device="cuda")
x = torch.randn(378, 3, 224, 224, device="cuda").half()
for _ in range(2):
    comp.compress(x)
torch.cuda.synchronize()
