"""GPU parity: fused EntropyBottleneck rate forward / backward vs the oracle (floating point:
rtol 1e-4 on likelihoods, 1e-5 on per-row rates, as north_star's 1e-4 relative bound)."""
import os

import numpy as np
import pytest
import torch

from oracle import eb as oeb

pytestmark = pytest.mark.gpu
TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]


def _load(weights_dir, tag, device):
    from lossyless_b200.entropy_models import EntropyBottleneck

    sd = torch.load(os.path.join(weights_dir, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)
    eb = EntropyBottleneck(512, init_scale=10, filters=[3, 3, 3, 3])
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items() if k.startswith("entropy_bottleneck.")})
    return eb.to(device), sd


@pytest.mark.parametrize("tag", TAGS)
def test_forward_golden(cuda_device, weights_dir, golden_dir, tag):
    from lossyless_b200.entropy_models import eb_rate_forward

    g = np.load(os.path.join(golden_dir, "eb_likelihood.npz"))
    eb, sd = _load(weights_dir, tag, cuda_device)
    z = torch.from_numpy(g[f"{tag}/z"]).to(cuda_device)
    b, s = sd["biasing"].to(cuda_device), sd["scaling"].to(cuda_device)
    with torch.no_grad():
        z_hat, q, rates = eb_rate_forward(eb, z, None, b, s)
    assert np.allclose(q.cpu().numpy(), g[f"{tag}/q_eval"], rtol=1e-4, atol=1e-10)
    ref_zhat = oeb.process_z_out(torch.from_numpy(g[f"{tag}/z_hat_in_eval"]), sd["biasing"], sd["scaling"])
    assert torch.allclose(z_hat.cpu(), ref_zhat, rtol=1e-6, atol=1e-6)
    ref_rates = -np.log(g[f"{tag}/q_eval"]).sum(-1)
    assert np.allclose(rates.cpu().numpy(), ref_rates, rtol=1e-5)
    noise = torch.from_numpy(g[f"{tag}/noise"]).to(cuda_device)
    with torch.no_grad():
        z_hat, q, rates = eb_rate_forward(eb, z, noise, b, s)
    assert np.allclose(q.cpu().numpy(), g[f"{tag}/q_train"], rtol=1e-4, atol=1e-10)
    assert np.allclose(rates.cpu().numpy(), -np.log(g[f"{tag}/q_train"]).sum(-1), rtol=1e-5)
    # compressai-layout module call ([B,C,1,1]) in eval mode
    eb.eval()
    out, lik = eb((oeb.process_z_in(z.cpu(), sd["biasing"], sd["scaling"])).to(cuda_device).reshape(-1, 512, 1, 1))
    assert out.shape == (16, 512, 1, 1)
    assert lik.requires_grad  # the likelihood is differentiable w.r.t. the bottleneck parameters, as compressai's
    assert np.allclose(lik.detach().reshape(16, 512).cpu().numpy(), g[f"{tag}/q_eval"], rtol=1e-4, atol=1e-10)


def test_backward_vs_oracle_autograd(cuda_device, weights_dir):
    """gradients of sum(w * rates) + sum(v * z_hat) wrt z, biasing, scaling and every bottleneck
    parameter, against torch autograd through the oracle restatement on the CPU."""
    from lossyless_b200.entropy_models import eb_rate_forward

    eb, sd = _load(weights_dir, "beta5e-02", cuda_device)
    gen = torch.Generator().manual_seed(0)
    B = 64
    z = torch.randn(B, 512, generator=gen) * 0.45
    noise = torch.rand(B, 512, generator=gen) - 0.5
    w = torch.rand(B, generator=gen) + 0.5
    v = torch.randn(B, 512, generator=gen) * 0.1

    # oracle side
    ebo = oeb.split_state_dict(sd)
    leaves = ebo["matrices"] + ebo["biases"] + ebo["factors"]
    for t in leaves:
        t.requires_grad_(True)
    zo = z.clone().requires_grad_(True)
    bo = sd["biasing"].clone().requires_grad_(True)
    so = sd["scaling"].clone().requires_grad_(True)
    zh_in, q = oeb.forward(ebo, oeb.process_z_in(zo, bo, so), noise)
    loss_o = (w * (-torch.log(q).sum(-1))).sum() + (v * oeb.process_z_out(zh_in, bo, so)).sum()
    loss_o.backward()

    # product side
    eb.train()
    zp = z.clone().to(cuda_device).requires_grad_(True)
    bp = sd["biasing"].clone().to(cuda_device).requires_grad_(True)
    sp = sd["scaling"].clone().to(cuda_device).requires_grad_(True)
    z_hat, _, rates = eb_rate_forward(eb, zp, noise.to(cuda_device), bp, sp)
    loss_p = (w.to(cuda_device) * rates).sum() + (v.to(cuda_device) * z_hat).sum()
    loss_p.backward()
    assert abs(loss_p.item() - loss_o.item()) < 1e-4 * abs(loss_o.item())

    def close(a, b, name, rtol=2e-3):
        a, b = a.detach().cpu().double(), b.detach().double()
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() < rtol * scale, (name, (a - b).abs().max().item(), scale)

    close(zp.grad, zo.grad, "z")
    close(bp.grad, bo.grad, "biasing")
    close(sp.grad, so.grad, "scaling")
    for i in range(5):
        close(getattr(eb, f"_matrix{i}").grad, ebo["matrices"][i].grad, f"matrix{i}")
        close(getattr(eb, f"_bias{i}").grad, ebo["biases"][i].grad, f"bias{i}")
        if i < 4:
            close(getattr(eb, f"_factor{i}").grad, ebo["factors"][i].grad, f"factor{i}")


def test_large_rows_and_shapes(cuda_device, weights_dir):
    """sweep-size input (65 536 x 512) stays finite and matches a 3-row oracle spot check;
    C = 128 (the aug-MNIST configuration's width) also works."""
    from lossyless_b200.entropy_models import EntropyBottleneck, eb_rate_forward

    eb, sd = _load(weights_dir, "beta1e-02", cuda_device)
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    z = torch.randn(65536, 512, generator=gen, device=cuda_device) * 0.45
    b, s = sd["biasing"].to(cuda_device), sd["scaling"].to(cuda_device)
    with torch.no_grad():
        z_hat, q, rates = eb_rate_forward(eb, z, None, b, s)
    assert torch.isfinite(rates).all() and (q >= 1e-9).all() and (q <= 1).all()
    ebo = oeb.split_state_dict(sd)
    rows = [0, 31337, 65535]
    _, q_ref = oeb.forward(ebo, oeb.process_z_in(z[rows].cpu(), sd["biasing"], sd["scaling"]))
    assert np.allclose(q[rows].cpu().numpy(), q_ref.numpy(), rtol=1e-4, atol=1e-10)
    small = EntropyBottleneck(128, init_scale=10, filters=[3, 3, 3, 3]).to(cuda_device).eval()
    x = torch.randn(33, 128, device=cuda_device) * 4
    with torch.no_grad():
        out, lik = small(x)
    ebs = oeb.split_state_dict({f"entropy_bottleneck.{k}": v.cpu() for k, v in small.state_dict().items()})
    _, q_ref = oeb.forward(ebs, x.cpu())
    assert np.allclose(lik.cpu().numpy(), q_ref.numpy(), rtol=1e-4, atol=1e-10)
