"""The first "next" row: scale/mean hyperprior rate -- compressai GaussianConditional +
reference HRateHyperprior (lossyless/rates.py:571-756).  CPU tests: oracle and host tables; GPU tests:
indexed rANS bytes, likelihood, round trips, and that the real bit-rate tracks the entropy estimate
(i.e. the reference's means/scales mix-up at rates.py:691-692 is not reproduced)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import eb as oeb
from oracle import gaussian as og

KW = dict(z_dim=512, kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "gaussian_conditional.npz"))


def test_oracle_tables_golden_and_invariants(golden_dir):
    g = _golden(golden_dir)
    tab = og.update(og.get_scale_table())
    for k in ("cdf", "cdf_length", "offset"):
        assert np.array_equal(tab[k], g[k])
    assert tab["cdf"].shape == (64, 3133) and tab["cdf_length"][0] == 5 and tab["offset"][-1] == -1565
    for i in range(64):
        row = tab["cdf"][i, : tab["cdf_length"][i]]
        assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) > 0).all()
    # scale table: 64 log-spaced levels 0.11 .. 256 (lossyless/rates.py:567)
    st = og.get_scale_table()
    assert len(st) == 64 and st[0].item() == pytest.approx(0.11) and st[-1].item() == pytest.approx(256, rel=1e-5)
    # oracle round trip with per-element indexes and means, escapes included
    x, means, idx = torch.from_numpy(g["x"]), torch.from_numpy(g["means"]), torch.from_numpy(g["indexes"])
    strings, sym = og.compress(x, idx, means, tab)
    assert np.array_equal(sym, g["symbols"]) and b"".join(strings) == g["bytes"].tobytes()
    _, sym2 = og.decompress(strings, idx, means, tab)
    assert np.array_equal(sym2, sym)


def test_product_host_tables_and_state_dict(native_lib, golden_dir):
    from lossyless_b200.rates import get_rate_estimator

    g = _golden(golden_dir)
    r = get_rate_estimator("H_hyper", **KW)
    assert r.side_z_dim == 102  # 512 // 5, hub/compressor.py:43
    assert not r.is_coder_updated
    r.update(force=True)
    gc = r.gaussian_conditional
    assert np.array_equal(gc._quantized_cdf.numpy(), g["cdf"]) and np.array_equal(gc._offset.numpy(), g["offset"])
    assert np.array_equal(gc.build_indexes(torch.from_numpy(g["scales"])).numpy(), g["indexes"])
    r2 = get_rate_estimator("H_hyper", **KW)
    r2.load_state_dict(r.state_dict())  # dynamically sized buffers (CDFs, scale_table) resize on load
    assert r2.is_coder_updated and torch.equal(r2.gaussian_conditional._quantized_cdf, gc._quantized_cdf)
    assert len(r.aux_parameters()) == 1  # only the side bottleneck's quantiles


@pytest.mark.gpu
def test_indexed_rans_bytes(cuda_device, golden_dir):
    from lossyless_b200.entropy_models import GaussianConditional

    g = _golden(golden_dir)
    gc = GaussianConditional(None)
    gc.update_scale_table(og.get_scale_table().tolist())
    gc = gc.to(cuda_device)
    x, means, idx = (torch.from_numpy(g[k]).to(cuda_device) for k in ("x", "means", "indexes"))
    strings = gc.compress(x, idx, means=means)
    lens = g["lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    gold = g["bytes"].tobytes()
    assert [len(s) for s in strings] == lens.tolist()
    assert all(s == gold[off[i]: off[i + 1]] for i, s in enumerate(strings))
    z_hat = gc.decompress(strings, idx, means=means)
    assert torch.equal(z_hat.cpu(), torch.from_numpy(g["symbols"]).float() + torch.from_numpy(g["means"]))
    # likelihood (plain torch ops on the GPU) against the oracle
    scales = torch.from_numpy(g["scales"]).to(cuda_device)
    gc.eval()
    out, lik = gc(x, scales, means=means)
    out_o, lik_o = og.forward(torch.from_numpy(g["x"]), torch.from_numpy(g["scales"]), torch.from_numpy(g["means"]))
    assert torch.allclose(out.cpu(), out_o) and np.allclose(lik.cpu().numpy(), lik_o.numpy(), rtol=1e-4, atol=1e-10)


def _oracle_hyper(r, z, noise=None):
    """the rate of HRateHyperprior.forward_help restated with the oracle pieces (CPU, fp32)."""
    cpu = lambda t: t.detach().float().cpu()
    b, s = cpu(r.biasing), cpu(r.scaling)
    z_in = oeb.process_z_in(z, b, s)

    def mlp(m, x):
        for layer in m.module:
            if isinstance(layer, torch.nn.Linear):
                x = x @ cpu(layer.weight).t() + cpu(layer.bias)
            elif isinstance(layer, torch.nn.ReLU):
                x = torch.relu(x)
        return x

    side = mlp(r.side_encoder, z_in)
    ebo = oeb.split_state_dict({f"entropy_bottleneck.{k}": cpu(v) for k, v in r.entropy_bottleneck.state_dict().items()})
    side_hat, q_s = oeb.forward(ebo, side, None if noise is None else noise[1])
    scales, means = mlp(r.z_encoder, side_hat).chunk(2, -1)
    z_hat_in, q_z = og.forward(z_in, scales, means, None if noise is None else noise[0])
    rates = -torch.log(q_s).sum(-1) - torch.log(q_z).sum(-1)
    return oeb.process_z_out(z_hat_in, b, s), rates


@pytest.mark.gpu
def test_hyperprior_rate_and_codec(cuda_device):
    from lossyless_b200.rates import get_rate_estimator

    torch.manual_seed(0)
    r = get_rate_estimator("H_hyper", **KW).to(cuda_device)
    gen = torch.Generator().manual_seed(3)
    z = torch.randn(24, 512, generator=gen) * 0.45
    with torch.no_grad():
        r.scaling.add_(1.5)
    # eval mode
    r.eval()
    with torch.no_grad():
        z_hat, rates, logs, _ = r(z.to(cuda_device), None, None)
    z_hat_o, rates_o = _oracle_hyper(r, z)
    assert torch.allclose(rates.cpu(), rates_o, rtol=1e-4) and torch.allclose(z_hat.cpu(), z_hat_o, atol=1e-5)
    # training mode with injected noise, gradients reach every trainable part
    r.train()
    noise = (torch.rand(24, 512, generator=gen) - 0.5, torch.rand(24, 102, generator=gen) - 0.5)
    zg = z.clone().to(cuda_device).requires_grad_(True)
    z_hat, rates, _, _ = r.forward_help(zg, None, None, noise=tuple(n.to(cuda_device) for n in noise))
    _, rates_o = _oracle_hyper(r, z, noise)
    assert torch.allclose(rates.detach().cpu(), rates_o, rtol=1e-4)
    rates.sum().backward()
    assert zg.grad is not None and r.scaling.grad is not None and r.entropy_bottleneck._matrix0.grad is not None
    assert r.side_encoder.module[0].weight.grad is not None and r.z_encoder.module[-1].weight.grad is not None
    # real coding: two streams per image, exact round trip, and the bit-rate tracks the entropy estimate
    r.eval()
    r.prepare_compressor_()
    with torch.no_grad():
        z_hat, rates, logs, _ = r(z.to(cuda_device), None, None)
        strings = r.compress(z.to(cuda_device))
        z_dec = r.decompress(strings)
    assert len(strings) == 2 and len(strings[0]) == 24 and len(strings[1]) == 24
    assert torch.allclose(z_dec, z_hat, atol=1e-5)
    n_bits, h = logs["n_bits"], float(logs["H_q_Z"])
    # (untrained model: escapes are cheaper than the 1e-9 likelihood floor and scales are quantised to 64
    # levels, so the match is loose; with the reference's means/scales mix-up n_bits would be several x H)
    assert 0.75 * h < n_bits < 1.25 * h + 2 * 64, (n_bits, h)
    for k in ("H_q_ZlS", "H_q_S", "H_q_Z", "n_bits", "compress_time", "receiver_time"):
        assert k in logs


@pytest.mark.gpu
def test_learnable_compressor_with_hyperprior(cuda_device):
    """README's CLIP recipe (config/featurizer/bottleneck_clip_lossyZ.yaml: rate=H_hyper, lossy_Z, frozen CLIP)."""
    from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams

    hp = default_clip_hparams(distortion="lossy_Z", beta=5e-2, max_steps=20, n_data=5000)
    hp.rate = type(hp)(dict(mode="H_hyper", factor_beta=1, kwargs=dict(z_dim=512, is_endToEnd=False, kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))))
    model = LearnableCompressor(hp).to(cuda_device)
    x = torch.randn(16, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    batch = (x, (None, x))
    model.train()
    first = float(model.step(batch)[0])
    logged = model.fit([batch] * 15)
    assert logged["train/feat/loss"] * math.log(2) < first
    assert "train/feat/H_q_S" in logged and "train/feat/coder_loss" in logged
    model.eval()
    model.on_test_epoch_start()
    with torch.no_grad():
        model.test_step(batch, 0)
    assert model.logged["test/feat/n_bits"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("with_means", [True, False])
def test_gaussian_conditional_fused_kernel_vs_autograd(cuda_device, training, with_means):
    """ll_gc_forward / ll_gc_backward against torch autograd through the oracle's op-by-op restatement of compressai's
    GaussianConditional.forward (both LowerBounds with compressai's gradient rule), gradients arriving through all three
    outputs (outputs, likelihood, per-row rate).  Scales below the 0.11 bound and likelihoods at the 1e-9 bound included."""
    from lossyless_b200.entropy_models import GaussianConditional

    gen = torch.Generator().manual_seed(3)
    B, C = 37, 64
    x = torch.randn(B, C, generator=gen) * 4
    x[0] *= 30  # far tails: likelihood clamps at 1e-9
    scales = torch.rand(B, C, generator=gen) * 3 + 0.02  # some below the 0.11 bound
    means = torch.randn(B, C, generator=gen) if with_means else None
    noise = torch.rand(B, C, generator=gen) - 0.5 if training else None
    w_out, w_q, w_r = torch.randn(B, C, generator=gen), torch.randn(B, C, generator=gen), torch.rand(B, generator=gen) + 0.5

    def reference(x, scales, means):
        out = x + noise if training else torch.round(x - (means if means is not None else 0)) + (means if means is not None else 0)
        v = torch.abs(out - means) if means is not None else torch.abs(out)
        s = oeb._LowerBound.apply(scales, og.SCALE_BOUND)
        lik = og.standardized_cumulative((0.5 - v) / s) - og.standardized_cumulative((-0.5 - v) / s)
        q = oeb._LowerBound.apply(lik, og.LIKELIHOOD_BOUND)
        return out, q, -torch.log(q).sum(-1)

    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (x, scales, means)]
    out_r, q_r, r_r = reference(*leaves)
    ((out_r * w_out).sum() + (q_r * w_q).sum() + (r_r * w_r).sum()).backward()

    gc = GaussianConditional(None).to(cuda_device)
    gc.train(training)
    dl = [t.clone().to(cuda_device).requires_grad_(True) if t is not None else None for t in (x, scales, means)]
    out, q, rates = gc.forward_rates(dl[0], dl[1], means=dl[2], noise=noise.to(cuda_device) if training else None)
    ((out * w_out.to(cuda_device)).sum() + (q * w_q.to(cuda_device)).sum() + (rates * w_r.to(cuda_device)).sum()).backward()
    assert torch.allclose(out.detach().cpu(), out_r.detach(), rtol=1e-6, atol=1e-6)
    assert np.allclose(q.detach().cpu().numpy(), q_r.detach().numpy(), rtol=1e-4, atol=1e-12)
    assert np.allclose(rates.detach().cpu().numpy(), r_r.detach().numpy(), rtol=1e-5)
    for got, want, name in zip(dl, leaves, ("x", "scales", "means")):
        if want is None:
            continue
        g, w = got.grad.cpu().double(), want.grad.double()
        assert (g - w).abs().max().item() < 2e-4 * (w.abs().max().item() + 1e-9), (name, (g - w).abs().max().item(), w.abs().max().item())
    # the module-level forward (compressai signature) returns the same outputs / likelihood
    out2, q2 = gc(dl[0].detach(), dl[1].detach(), means=dl[2].detach() if dl[2] is not None else None, noise=noise.to(cuda_device) if training else None)
    assert torch.equal(out2, out.detach()) and torch.equal(q2, q.detach())
