"""Generates the committed golden fixtures in tests/golden/ from the CPU oracle.

The real reference cannot be imported here (compressai / clip / pytorch_lightning are not
installed and there is no network), so these vectors come from the
oracle restatement run on the published checkpoints -- they pin the oracle against drift and
let the GPU box (which has no /root/reference) check the CUDA path against fixed bytes.
PARITY UNPINNED with respect to compressai itself.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import eb as oeb  # noqa: E402
from oracle import losses as olosses  # noqa: E402
from oracle import rans  # noqa: E402
from oracle import vit as ovit  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]
WEIGHTS = os.path.join(ROOT, "lossyless_b200", "hub", "weights")


def load(tag):
    return torch.load(os.path.join(WEIGHTS, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)


def main():
    tables, vectors, lik = {}, {}, {}
    for tag in TAGS:
        sd = load(tag)
        eb = oeb.split_state_dict(sd)
        tab = oeb.update(eb)
        for k in ("cdf", "cdf_length", "offset"):
            tables[f"{tag}/{k}"] = tab[k]

        # --- codec vectors: CLIP-like z, plus rows that force escapes of every nibble count
        g = torch.Generator().manual_seed(1234)
        z = (torch.randn(24, 512, generator=g) * 0.45).half()
        z[16] *= 8.0  # many small escapes
        z[17] = (torch.randn(512, generator=g) * 40).half()  # multi-nibble escapes both signs
        z[18] = 0.0
        z[19, ::2], z[19, 1::2] = 60000.0, -60000.0  # fp16 extremes -> 6-nibble raw values
        z[20] = torch.linspace(-3, 3, 512).half()
        z_in = oeb.process_z_in(z, sd["biasing"], sd["scaling"])
        strings, sym = oeb.compress(eb, tab, z_in)
        z_hat_in, sym2 = oeb.decompress(eb, tab, strings)
        assert (sym == sym2).all()
        z_hat = oeb.process_z_out(z_hat_in, sd["biasing"], sd["scaling"])
        vectors[f"{tag}/z_f16"] = z.numpy()
        vectors[f"{tag}/symbols"] = sym
        vectors[f"{tag}/lens"] = np.array([len(s) for s in strings], dtype=np.int32)
        vectors[f"{tag}/bytes"] = np.frombuffer(b"".join(strings), dtype=np.uint8)
        vectors[f"{tag}/z_hat"] = z_hat.numpy()

        # --- likelihood vectors (eval mode and training mode with injected noise) + aux loss
        g = torch.Generator().manual_seed(99)
        zz = torch.randn(16, 512, generator=g) * 0.45
        noise = torch.rand(16, 512, generator=g) - 0.5
        zin = oeb.process_z_in(zz, sd["biasing"], sd["scaling"])
        zh_e, q_e = oeb.forward(eb, zin)
        zh_t, q_t = oeb.forward(eb, zin, noise)
        lik[f"{tag}/z"] = zz.numpy()
        lik[f"{tag}/noise"] = noise.numpy()
        lik[f"{tag}/z_hat_in_eval"] = zh_e.numpy()
        lik[f"{tag}/q_eval"] = q_e.numpy()
        lik[f"{tag}/z_hat_in_train"] = zh_t.numpy()
        lik[f"{tag}/q_train"] = q_t.numpy()
        lik[f"{tag}/aux_loss"] = np.array(oeb.aux_loss(eb, sd["entropy_bottleneck.target"]).item(), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "cdf_tables.npz"), **tables)
    np.savez_compressed(os.path.join(OUT, "rans_vectors.npz"), **vectors)
    np.savez_compressed(os.path.join(OUT, "eb_likelihood.npz"), **lik)

    # --- ViT-B/32 fp32 oracle output on seeded weights/inputs (weights are re-derived from the
    # seed by oracle.vit.init_state_dict; only x and z are stored)
    sdv = ovit.init_state_dict(0)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, 224, 224, generator=g).half()
    with torch.no_grad():
        z = ovit.forward(sdv, x.float())
    np.savez_compressed(os.path.join(OUT, "vit_b32.npz"), x_f16=x.numpy(), z=z.numpy())

    # --- contrastive / lossy-Z / loss vectors (W = 1 and the W = 2 gather quirk)
    g = torch.Generator().manual_seed(5)
    B = 8
    proj = dict(w1=torch.randn(128, 512, generator=g) * 0.05, b1=torch.randn(128, generator=g) * 0.01,
                w2=torch.randn(256, 128, generator=g) * 0.1, b2=torch.randn(256, generator=g) * 0.01)
    zs = [torch.randn(B, 512, generator=g) for _ in range(4)]
    tile0 = olosses.contrastive_local_logits(zs[0], zs[1], proj)
    tile1 = olosses.contrastive_local_logits(zs[2], zs[3], proj)
    ls = torch.log(torch.tensor(1 / 0.07))
    d1, logs1 = olosses.contrastive_distortion(zs[0], zs[1], proj, temperature=0.01, logit_scale=ls, effective_batch_size=5000)
    d2r0, _ = olosses.contrastive_distortion(zs[0], zs[1], proj, rank=0, other_tiles=[tile1], temperature=0.01, logit_scale=ls, effective_batch_size=5000)
    d2r1, _ = olosses.contrastive_distortion(zs[2], zs[3], proj, rank=1, other_tiles=[tile0], temperature=0.01, logit_scale=ls, effective_batch_size=5000)
    lz = olosses.lossy_z_distortion(zs[0], zs[1])
    np.savez_compressed(
        os.path.join(OUT, "losses.npz"),
        **{f"proj_{k}": v.numpy() for k, v in proj.items()},
        **{f"z{i}": z.numpy() for i, z in enumerate(zs)},
        tile0=tile0.numpy(), contrastive_w1=d1.numpy(), I_q_zm_w1=np.array(float(logs1["I_q_zm"])),
        contrastive_w2_rank0=d2r0.numpy(), contrastive_w2_rank1=d2r1.numpy(), lossy_z=lz.numpy())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def gaussian_golden():
    """GaussianConditional tables for the reference's 64-level scale table + an indexed-rANS vector."""
    from oracle import gaussian as og

    tab = og.update(og.get_scale_table())
    g = torch.Generator().manual_seed(21)
    x = torch.randn(6, 512, generator=g) * 6
    means = torch.randn(6, 512, generator=g)
    scales = torch.rand(6, 512, generator=g) * 5 + 0.02
    x[4] *= 40  # escapes far outside the narrow tables
    scales[5] = 200.0  # widest tables (3133 entries)
    idx = og.build_indexes(scales, og.get_scale_table())
    strings, sym = og.compress(x, idx, means, tab)
    np.savez_compressed(os.path.join(OUT, "gaussian_conditional.npz"), cdf=tab["cdf"], cdf_length=tab["cdf_length"], offset=tab["offset"],
                        x=x.numpy(), means=means.numpy(), scales=scales.numpy(), indexes=idx.numpy(), symbols=sym,
                        lens=np.array([len(s) for s in strings], dtype=np.int32), bytes=np.frombuffer(b"".join(strings), dtype=np.uint8))


if __name__ == "__main__":
    gaussian_golden()
    main()
