"""CPU: the oracle against (a) the reference's few reproducible known answers, (b) its own
committed golden vectors, (c) an independent pure-Python restatement, (d) HF's CLIP
architecture, (e) codec invariants (round trips incl. escapes, CDF monotonicity)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import eb as oeb
from oracle import losses as olosses
from oracle import rans
from oracle import vit as ovit

TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]


def _sd(weights_dir, tag):
    return torch.load(os.path.join(weights_dir, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)


@pytest.mark.parametrize("tag", TAGS)
def test_checkpoint_known_answers(weights_dir, tag):
    """pins that need no data -- target = +-log(2/1e-9 - 1), 31 232 EB parameters,
    32 260 floats per checkpoint, empty CDF buffers."""
    sd = _sd(weights_dir, tag)
    t = math.log(2 / 1e-9 - 1)
    assert torch.allclose(sd["entropy_bottleneck.target"], torch.tensor([-t, 0.0, t]), atol=1e-4)
    n_eb = sum(v.numel() for k, v in sd.items() if k.startswith("entropy_bottleneck._") and v.dtype == torch.float32)
    assert n_eb + sd["entropy_bottleneck.quantiles"].numel() == 31232
    assert sum(v.numel() for v in sd.values() if v.dtype == torch.float32) == 32260
    for k in ("_offset", "_quantized_cdf", "_cdf_length"):
        assert sd[f"entropy_bottleneck.{k}"].numel() == 0
    # the trained medians sit where the aux loss puts them: logits(median) ~ 0 (the outer
    # quantiles are not converged in the published checkpoints, so only the median is a pin)
    eb = oeb.split_state_dict(sd)
    err = (oeb.logits_cumulative(eb, eb["quantiles"]) - sd["entropy_bottleneck.target"]).abs()
    assert err[:, 0, 1].median() < 0.05


def test_banana_known_answer_64_bits():
    """results/exp_banana_viz_VIC/summarized_metrics_featurizer.csv: n_bits == 64.0 for z_dim = 2,
    i.e. a 2-symbol message is exactly the bare rANS flush (two uint32 words, 64-bit state)."""
    cdf = np.array([[0, 20000, 45000, 65535, 65536]] * 2, dtype=np.int32)
    sizes, offsets = np.array([5, 5], np.int32), np.array([-1, -1], np.int32)
    for sym in ([0, 0], [-1, 1], [1, -1]):
        b = rans.encode_with_indexes(sym, [0, 1], cdf, sizes, offsets)
        assert len(b) * 8 == 64
        assert rans.decode_with_indexes(b, [0, 1], cdf, sizes, offsets).tolist() == sym


@pytest.mark.parametrize("tag", TAGS)
def test_tables_match_golden_and_invariants(weights_dir, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "cdf_tables.npz"))
    tab = oeb.update(oeb.split_state_dict(_sd(weights_dir, tag)))
    for k in ("cdf", "cdf_length", "offset"):
        assert np.array_equal(tab[k], g[f"{tag}/{k}"]), k
    for c in range(tab["cdf"].shape[0]):
        L = tab["cdf_length"][c]
        row = tab["cdf"][c, :L]
        assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) > 0).all()
    # table widths 32 / 32 / 33 and offsets in [-19, -8]
    assert tab["cdf"].shape[1] == {"beta1e-01": 32, "beta5e-02": 32, "beta1e-02": 33}[tag]
    assert tab["offset"].min() >= -19 and tab["offset"].max() <= -8


@pytest.mark.parametrize("tag", TAGS)
def test_pmf_float32_against_float64(weights_dir, tag):
    """floating point part of update(): float32 torch op order vs float64 numpy, tolerance 1e-6."""
    eb = oeb.split_state_dict(_sd(weights_dir, tag))
    tab = oeb.update(eb)
    L = tab["pmf"].shape[1]
    med = oeb.medians(eb).numpy().astype(np.float64)
    samples = np.arange(L)[None, :] + (med + tab["offset"])[:, None]
    p64 = oeb.likelihood_f64(eb, samples)
    assert np.abs(p64 - tab["pmf"]).max() < 1e-6


@pytest.mark.parametrize("tag", TAGS)
def test_rans_golden_roundtrip_and_python_twin(weights_dir, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "rans_vectors.npz"))
    sd = _sd(weights_dir, tag)
    eb = oeb.split_state_dict(sd)
    tab = oeb.update(eb)
    z = torch.from_numpy(g[f"{tag}/z_f16"])
    strings, sym = oeb.compress(eb, tab, oeb.process_z_in(z, sd["biasing"], sd["scaling"]))
    assert np.array_equal(sym, g[f"{tag}/symbols"])
    assert np.array_equal(np.array([len(s) for s in strings], np.int32), g[f"{tag}/lens"])
    assert b"".join(strings) == g[f"{tag}/bytes"].tobytes()
    # C quantiser == torch quantiser
    sym_c = rans.quantize_symbols(z.float().numpy(), sd["biasing"].numpy(), sd["scaling"].exp().numpy(), oeb.medians(eb).numpy())
    assert np.array_equal(sym_c, sym)
    # decode inverts encode, escapes of every size included
    z_hat_in, sym2 = oeb.decompress(eb, tab, strings)
    assert np.array_equal(sym2, sym)
    assert np.allclose(oeb.process_z_out(z_hat_in, sd["biasing"], sd["scaling"]).numpy(), g[f"{tag}/z_hat"])
    n_escape = ((sym - tab["offset"][None]) < 0) | ((sym - tab["offset"][None]) >= (tab["cdf_length"][None] - 2))
    assert n_escape[19].all() and n_escape[17].sum() > 200  # the fixture really exercises the bypass path
    # independent big-int Python implementation agrees byte for byte on a few rows
    idx = list(range(512))
    for r in (0, 16, 17, 19):
        b = rans.py_encode(sym[r].tolist(), idx, tab["cdf"], tab["cdf_length"], tab["offset"])
        assert b == strings[r]
        assert rans.py_decode(b, idx, tab["cdf"], tab["cdf_length"], tab["offset"]) == sym[r].tolist()
    for s in strings:
        assert len(s) % 4 == 0 and len(s) >= 8


def test_rans_edge_cases():
    """empty message, single symbol, table with the minimum 2 entries + sentinel, huge escapes."""
    cdf = np.array([[0, 65535, 65536, 0], [0, 1, 2, 65536]], dtype=np.int32)
    sizes, offsets = np.array([3, 4], np.int32), np.array([0, -1], np.int32)
    assert len(rans.encode_with_indexes([], [], cdf, sizes, offsets)) == 8  # bare flush
    for sym, idx in (([0], [0]), ([5], [0]), ([-7], [1]), ([2 ** 27, -(2 ** 27), 0, 1], [0, 1, 0, 1])):
        b = rans.encode_with_indexes(sym, idx, cdf, sizes, offsets)
        assert rans.decode_with_indexes(b, idx, cdf, sizes, offsets).tolist() == sym
        assert b == rans.py_encode(sym, idx, cdf, sizes, offsets)


def test_pmf_to_quantized_cdf_stealing():
    """zero-probability bins must end up with frequency >= 1, taken from the smallest donor > 1."""
    cdf = rans.pmf_to_quantized_cdf(np.array([0.5, 0.0, 0.25, 0.0, 0.25 - 1e-9, 1e-9], np.float32))
    assert cdf[0] == 0 and cdf[-1] == 65536 and (np.diff(cdf) > 0).all()
    with pytest.raises(ValueError):
        rans.pmf_to_quantized_cdf(np.array([0.5, -0.1], np.float32))


@pytest.mark.parametrize("tag", TAGS)
def test_likelihood_golden(weights_dir, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "eb_likelihood.npz"))
    sd = _sd(weights_dir, tag)
    eb = oeb.split_state_dict(sd)
    zin = oeb.process_z_in(torch.from_numpy(g[f"{tag}/z"]), sd["biasing"], sd["scaling"])
    zh, q = oeb.forward(eb, zin)
    assert np.allclose(q.numpy(), g[f"{tag}/q_eval"], rtol=1e-5, atol=1e-9)
    zh, q = oeb.forward(eb, zin, torch.from_numpy(g[f"{tag}/noise"]))
    assert np.allclose(q.numpy(), g[f"{tag}/q_train"], rtol=1e-5, atol=1e-9)
    assert (q.numpy() >= 1e-9).all() and (q.numpy() <= 1).all()
    assert abs(oeb.aux_loss(eb, sd["entropy_bottleneck.target"]).item() - float(g[f"{tag}/aux_loss"])) < 1e-2


def test_vit_oracle_matches_hf_clip_architecture():
    """independent architecture check: transformers' CLIPVisionModelWithProjection with the same
    (random) weights gives the same embedding (2 layers to keep the CPU suite short)."""
    transformers = pytest.importorskip("transformers")
    cfg = ovit.vit_config(layers=2)
    sd = ovit.init_state_dict(0, cfg)
    hc = transformers.CLIPVisionConfig(hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                                       num_attention_heads=12, image_size=224, patch_size=32,
                                       projection_dim=512, hidden_act="quick_gelu")
    m = transformers.CLIPVisionModelWithProjection(hc).eval()
    m.load_state_dict(ovit.to_hf_state_dict(sd, cfg))
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a, b = m(pixel_values=x).image_embeds, ovit.forward(sd, x, cfg)
    assert (a - b).abs().max() < 1e-4
    assert sum(v.numel() for v in ovit.init_state_dict(0).values()) == 87849216


def test_vit_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "vit_b32.npz"))
    with torch.no_grad():
        z = ovit.forward(ovit.init_state_dict(0), torch.from_numpy(g["x_f16"]).float())
    assert np.allclose(z.numpy(), g["z"], atol=2e-4)


def test_losses_golden_and_gather_quirk(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    t = lambda k: torch.from_numpy(g[k])
    proj = {k: t(f"proj_{k}") for k in ("w1", "b1", "w2", "b2")}
    ls = torch.log(torch.tensor(1 / 0.07))
    kw = dict(temperature=0.01, logit_scale=ls, effective_batch_size=5000)
    d1, logs = olosses.contrastive_distortion(t("z0"), t("z1"), proj, **kw)
    assert np.allclose(d1.numpy(), g["contrastive_w1"], rtol=1e-5)
    tile1 = olosses.contrastive_local_logits(t("z2"), t("z3"), proj)
    d2, logs2 = olosses.contrastive_distortion(t("z0"), t("z1"), proj, rank=0, other_tiles=[tile1], **kw)
    assert np.allclose(d2.numpy(), g["contrastive_w2_rank0"], rtol=1e-5)
    assert logs2["n_negatives"] == 2 * 16 - 1  # [2B, 2B*W] minus the own diagonal
    assert np.allclose(olosses.lossy_z_distortion(t("z0"), t("z1")).numpy(), g["lossy_z"], rtol=1e-6)
    # PairwiseDistance semantics: eps is added to the difference
    assert torch.allclose(olosses.lossy_z_distortion(t("z0"), t("z1")), torch.nn.PairwiseDistance(p=1)(t("z0"), t("z1")))


def test_annealer_and_rd_loss():
    a = olosses.Annealer(1e-5 * 0.01, 0.01, n_steps_anneal=100, mode="linear")
    assert a(0) == pytest.approx(1e-7) and a(100) == 0.01 and a(50) == pytest.approx(1e-7 + (0.01 - 1e-7) / 2)
    gm = olosses.Annealer(1e-7, 0.01, 100, mode="geometric")
    assert gm(50) == pytest.approx(math.sqrt(1e-7 * 0.01), rel=1e-6)
    rates = torch.tensor([3.0, 5.0], requires_grad=True)
    dist = torch.tensor([1.0, 2.0], requires_grad=True)
    loss, logs = olosses.rd_loss(rates, dist, curr_beta=0.001, final_beta=0.01, labda=2.0)
    assert loss.item() == pytest.approx(2.0 * 1.5 + 0.01 * 4.0)  # value uses final beta
    loss.backward()
    assert torch.allclose(rates.grad, torch.full((2,), 0.001 / 2))  # gradient uses the annealed beta
