"""One-off: re-serialise the reference's three published bottleneck checkpoints
(/root/reference/hub/beta*/factorized_rate.pt, the files hubconf.py:15 downloads from the
v1.0 GitHub release as beta{beta:0.0e}_factorized_rate.pt) into
lossyless_b200/hub/weights/ so the hub entry points work offline.  Weights are data, not
source; tensors and key names are preserved bit-for-bit (checked below)."""
import os
import sys

import torch

SRC = "/root/reference/hub"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lossyless_b200", "hub", "weights")

for tag in ["beta1e-01", "beta5e-02", "beta1e-02"]:
    sd = torch.load(os.path.join(SRC, tag, "factorized_rate.pt"), map_location="cpu", weights_only=True)
    out = os.path.join(DST, f"{tag}_factorized_rate.pt")
    torch.save({k: v.clone() for k, v in sd.items()}, out)
    back = torch.load(out, map_location="cpu", weights_only=True)
    assert list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    print(tag, {k: tuple(v.shape) for k, v in sd.items()}, file=sys.stderr)
    print("wrote", out, os.path.getsize(out))
