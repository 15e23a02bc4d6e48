"""GPU parity: contrastive logits / InfoNCE / lossy-Z kernels (forward and backward) against
torch autograd through the oracle restatement.  Floating point: rtol 1e-4 (north_star)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import losses as olosses

pytestmark = pytest.mark.gpu


def _proj_from_golden(g):
    return {k: torch.from_numpy(g[f"proj_{k}"]) for k in ("w1", "b1", "w2", "b2")}


def _module(proj, device, n_data=5000):
    from lossyless_b200.distortions import ContrastiveDistortion

    m = ContrastiveDistortion(temperature=0.01, is_train_temperature=True, effective_batch_size=n_data,
                              project_kwargs=dict(mode="mlp", out_shape=0.5, in_shape=512))
    lin1, lin2 = m.projector.module[0], m.projector.module[-1]
    assert lin1.weight.shape == (128, 512) and lin2.weight.shape == (256, 128)  # FlattenMLP(512->128->ReLU->256)
    with torch.no_grad():
        lin1.weight.copy_(proj["w1"]), lin1.bias.copy_(proj["b1"]), lin2.weight.copy_(proj["w2"]), lin2.bias.copy_(proj["b2"])
    return m.to(device)


def test_contrastive_golden_w1(cuda_device, golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    proj = _proj_from_golden(g)
    m = _module(proj, cuda_device)
    m.is_already_featurized = True  # feed z_pos_hat directly (the encoder pass is tested elsewhere)
    z0, z1 = torch.from_numpy(g["z0"]).to(cuda_device), torch.from_numpy(g["z1"]).to(cuda_device)
    d, logs, _ = m(z0, z1, None, None)
    assert np.allclose(d.detach().cpu().numpy(), g["contrastive_w1"], rtol=1e-4)
    assert float(logs["I_q_zm"]) == pytest.approx(float(g["I_q_zm_w1"]), rel=1e-4)
    assert logs["n_negatives"] == 2 * 8 - 1
    # the logits tile itself
    from lossyless_b200.distortions import _LogitsTile

    zs = torch.cat([m.projector(z0), m.projector(z1)])
    tile = _LogitsTile.apply(zs, 16)
    assert np.allclose(tile.detach().cpu().numpy(), g["tile0"], atol=2e-6)


@pytest.mark.parametrize("B,world", [(8, 1), (64, 1), (8, 2), (256, 1)])
def test_contrastive_forward_backward_vs_oracle(cuda_device, B, world):
    """loss and gradients wrt z_hat, z_pos_hat, projector weights and logit_scale; world=2 emulates
    the gathered remote tile (constant columns to the right of the own tile)."""
    from lossyless_b200.distortions import _LogitsTile

    gen = torch.Generator().manual_seed(B + world)
    proj = dict(w1=torch.randn(128, 512, generator=gen) * 0.05, b1=torch.randn(128, generator=gen) * 0.01,
                w2=torch.randn(256, 128, generator=gen) * 0.1, b2=torch.randn(256, generator=gen) * 0.01)
    z, zp = torch.randn(B, 512, generator=gen), torch.randn(B, 512, generator=gen)
    remote = olosses.contrastive_local_logits(torch.randn(B, 512, generator=gen), torch.randn(B, 512, generator=gen), proj) if world == 2 else None
    w = torch.rand(B, generator=gen) + 0.5
    ls0 = math.log(1 / 0.07)

    # oracle (CPU autograd)
    po = {k: v.clone().requires_grad_(True) for k, v in proj.items()}
    zo, zpo = z.clone().requires_grad_(True), zp.clone().requires_grad_(True)
    lso = torch.tensor(ls0, requires_grad=True)
    d_o, _ = olosses.contrastive_distortion(zo, zpo, po, rank=0, other_tiles=[remote] if world == 2 else None,
                                            temperature=0.01, logit_scale=lso, effective_batch_size=5000)
    (w * d_o).sum().backward()

    # product
    m = _module(proj, cuda_device)
    m.is_already_featurized = True
    zg, zpg = z.clone().to(cuda_device).requires_grad_(True), zp.clone().to(cuda_device).requires_grad_(True)
    if world == 1:
        d_p, logs, _ = m(zg, zpg, None, None)
    else:
        zs = torch.cat([m.projector(zg), m.projector(zpg)])
        logits = torch.cat([_LogitsTile.apply(zs, 2 * B), remote.to(cuda_device)], dim=1)
        h, logs = m.compute_loss(logits)
        d_p = (h[:B] + h[B:]) / 2
        assert logs["n_negatives"] == 4 * B - 1
    (w.to(cuda_device) * d_p).sum().backward()

    assert torch.allclose(d_p.detach().cpu(), d_o.detach(), rtol=1e-4, atol=1e-5)

    def close(a, b, name, rtol=2e-3):
        a, b = a.detach().cpu().double(), b.detach().double()
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() < rtol * scale, (name, (a - b).abs().max().item(), scale)

    close(zg.grad, zo.grad, "z_hat")
    close(zpg.grad, zpo.grad, "z_pos_hat")
    close(m.projector.module[0].weight.grad, po["w1"].grad, "w1")
    close(m.projector.module[-1].weight.grad, po["w2"].grad, "w2")
    close(m.projector.module[-1].bias.grad, po["b2"].grad, "b2")
    close(m.logit_scale.grad, lso.grad, "logit_scale")


def test_temperature_clamp_stops_gradient(cuda_device):
    """logit_scale above log(1/temperature) is clamped: no gradient flows to it (CLIP scheme)."""
    gen = torch.Generator().manual_seed(0)
    proj = dict(w1=torch.randn(128, 512, generator=gen) * 0.05, b1=torch.zeros(128), w2=torch.randn(256, 128, generator=gen) * 0.1, b2=torch.zeros(256))
    m = _module(proj, cuda_device)
    m.is_already_featurized = True
    with torch.no_grad():
        m.logit_scale.fill_(math.log(1 / 0.01) + 1.0)
    z, zp = torch.randn(8, 512, generator=gen).cuda(), torch.randn(8, 512, generator=gen).cuda()
    d, _, _ = m(z, zp, None, None)
    d.sum().backward()
    assert float(m.logit_scale.grad) == 0.0


def test_lossy_z(cuda_device, golden_dir):
    from lossyless_b200.distortions import LossyZDistortion
    from lossyless_b200.distributions import Deterministic

    g = np.load(os.path.join(golden_dir, "losses.npz"))
    z_hat = torch.from_numpy(g["z0"]).to(cuda_device).requires_grad_(True)
    z = torch.from_numpy(g["z1"]).to(cuda_device)
    d, _, _ = LossyZDistortion(p_norm=1)(z_hat, None, Deterministic(z), None)
    assert np.allclose(d.detach().cpu().numpy(), g["lossy_z"], rtol=1e-6)
    d.sum().backward()
    assert torch.equal(z_hat.grad, torch.sign(z_hat.detach() - z + 1e-6))
