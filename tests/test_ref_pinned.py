"""CPU: the oracle (and the package's pure-host logic) against fixtures that were produced by EXECUTING THE
REFERENCE'S OWN PYTHON (tests/golden/make_golden_ref.py: /root/reference/lossyless and /root/reference/hub run in
place, with stand-ins only for the un-installable compressai / clip leaves).  This is what ties the oracle's
restatements of lossyless/distortions.py, helpers.py (Annealer, GatherFromGpus), learnable_compressors.py (loss,
step), rates.py (HRateFactorizedPrior / HRateHyperprior wrappers) and hub/compressor.py (container writer) to the
reference's code instead of to the builder's reading of it.  Tolerances: integer / byte work bit-exact, floating
point 1e-5 relative (same float32 torch ops, possibly in another association order)."""
import math
import os
import struct

import numpy as np
import pytest
import torch

from oracle import eb as oeb
from oracle import gaussian as og
from oracle import losses as olosses

TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]


@pytest.fixture(scope="module")
def ref(golden_dir):
    return {n: np.load(os.path.join(golden_dir, f"ref_{n}.npz")) for n in ("losses", "rates", "step", "hub")}


def T(a):
    return torch.from_numpy(np.asarray(a))


def _proj(g, prefix="contrastive/"):
    return dict(w1=T(g[prefix + "projector.module.0.weight"]), b1=T(g[prefix + "projector.module.0.bias"]),
                w2=T(g[prefix + "projector.module.4.weight"]), b2=T(g[prefix + "projector.module.4.bias"]))


# ------------------------------------------------------------------------------------------ distortions
def test_contrastive_w1_matches_reference(ref):
    g = ref["losses"]
    proj = {k: v.clone().requires_grad_(True) for k, v in _proj(g).items()}
    z, zp = T(g["w1/z_hat"]).requires_grad_(True), T(g["w1/z_pos"]).requires_grad_(True)
    ls = T(g["contrastive/logit_scale"]).clone().requires_grad_(True)
    tile = olosses.contrastive_local_logits(z, zp, proj)
    assert torch.allclose(tile, T(g["w1/logits"]), atol=1e-6)
    h, _ = olosses.contrastive_loss(tile, temperature=0.01, logit_scale=ls, effective_batch_size=5000)
    assert torch.allclose(h, T(g["w1/hat_H_mlz"]), rtol=1e-5)
    d, logs = olosses.contrastive_distortion(z, zp, proj, temperature=0.01, logit_scale=ls, effective_batch_size=5000)
    assert torch.allclose(d, T(g["w1/distortion"]), rtol=1e-5)
    assert float(logs["I_q_zm"]) == pytest.approx(float(g["w1/I_q_zm"]), rel=1e-5)
    assert float(logs["hat_H_m"]) == pytest.approx(float(g["w1/hat_H_m"]), rel=1e-6)
    assert int(logs["n_negatives"]) == int(g["w1/n_negatives"]) == 15
    d.mean().backward()
    for got, want in ((z.grad, "g_z_hat"), (zp.grad, "g_z_pos"), (ls.grad, "g_logit_scale"), (proj["w1"].grad, "g_w1"), (proj["w2"].grad, "g_w2")):
        w = T(g["w1/" + want])
        assert (got - w).abs().max() <= 1e-5 * w.abs().max() + 1e-9, want


def test_contrastive_plain_flavour_matches_reference(ref):
    """fixed temperature, no effective batch size, no projector, dot-product logits (distortions.py:322,378-381)."""
    g = ref["losses"]
    z, zp = T(g["plain/z_hat"]), T(g["plain/z_pos"])
    tile = olosses.contrastive_local_logits(z, zp, None, is_cosine=False)
    h, logs = olosses.contrastive_loss(tile, temperature=0.5)
    assert torch.allclose((h[:6] + h[6:]) / 2, T(g["plain/distortion"]), rtol=1e-5)
    assert float(logs["I_q_zm"]) == pytest.approx(float(g["plain/I_q_zm"]), rel=1e-5)
    assert int(logs["n_negatives"]) == int(g["plain/n_negatives"])


@pytest.mark.parametrize("rank", [0, 1])
def test_contrastive_w2_gather_quirk_matches_reference(ref, rank):
    """Two real gloo processes ran lossyless/helpers.py GatherFromGpus: own tile first, the other rank's
    SELF-similarity tile to its right, gradient through the own tile only."""
    g = ref["losses"]
    proj = {k: v.clone().requires_grad_(True) for k, v in _proj(g).items()}
    ls = T(g["contrastive/logit_scale"]).clone().requires_grad_(True)
    z = {r: T(g[f"w2/rank{r}/z_hat"]) for r in (0, 1)}
    zp = {r: T(g[f"w2/rank{r}/z_pos"]) for r in (0, 1)}
    other = 1 - rank
    with torch.no_grad():
        remote = olosses.contrastive_local_logits(z[other], zp[other], {k: v.detach() for k, v in proj.items()})
    zr, zpr = z[rank].clone().requires_grad_(True), zp[rank].clone().requires_grad_(True)
    tile = olosses.contrastive_local_logits(zr, zpr, proj)
    tiles = [None, None]
    tiles[rank], tiles[other] = tile, remote
    logits = olosses.contrastive_gather(tiles, rank)
    assert logits.shape == (16, 32)
    assert torch.allclose(logits, T(g[f"w2/rank{rank}/logits"]), atol=1e-6)
    d, logs = olosses.contrastive_distortion(zr, zpr, proj, rank=rank, other_tiles=[remote], temperature=0.01, logit_scale=ls,
                                             effective_batch_size=5000)
    assert torch.allclose(d, T(g[f"w2/rank{rank}/distortion"]), rtol=1e-5)
    assert int(logs["n_negatives"]) == int(g[f"w2/rank{rank}/n_negatives"]) == 31
    assert float(logs["I_q_zm"]) == pytest.approx(float(g[f"w2/rank{rank}/I_q_zm"]), rel=1e-5)
    d.mean().backward()
    for got, want in ((zr.grad, "g_z_hat"), (zpr.grad, "g_z_pos"), (ls.grad, "g_logit_scale"), (proj["w1"].grad, "g_w1")):
        w = T(g[f"w2/rank{rank}/{want}"])
        assert (got - w).abs().max() <= 1e-5 * w.abs().max() + 1e-9, want


@pytest.mark.parametrize("rank", [0, 3, 7])
def test_contrastive_w8_matches_reference(ref, golden_dir, rank):
    """SURVEY 8(d) configuration 3 runs on 8 ranks: the reference's ContrastiveDistortion over EIGHT real gloo processes
    (ref_losses_w8.npz, 4 samples per rank).  Own tile first, the seven other ranks' self-similarity tiles in rank order
    to its right, 63 negatives, gradient through the own tile only."""
    g8 = np.load(os.path.join(golden_dir, "ref_losses_w8.npz"))
    g = ref["losses"]
    proj = {k: v.clone().requires_grad_(True) for k, v in _proj(g).items()}
    ls = T(g["contrastive/logit_scale"]).clone().requires_grad_(True)
    frozen = {k: v.detach() for k, v in proj.items()}
    with torch.no_grad():
        remotes = [olosses.contrastive_local_logits(T(g8[f"w8/rank{r}/z_hat"]), T(g8[f"w8/rank{r}/z_pos"]), frozen)
                   for r in range(8) if r != rank]
    zr = T(g8[f"w8/rank{rank}/z_hat"]).clone().requires_grad_(True)
    zpr = T(g8[f"w8/rank{rank}/z_pos"]).clone().requires_grad_(True)
    tiles = list(remotes)
    tiles.insert(rank, olosses.contrastive_local_logits(zr, zpr, proj))
    logits = olosses.contrastive_gather(tiles, rank)
    assert logits.shape == (8, 64)
    assert torch.allclose(logits, T(g8[f"w8/rank{rank}/logits"]), atol=1e-6)
    d, logs = olosses.contrastive_distortion(zr, zpr, proj, rank=rank, other_tiles=remotes, temperature=0.01, logit_scale=ls,
                                             effective_batch_size=5000)
    assert torch.allclose(d, T(g8[f"w8/rank{rank}/distortion"]), rtol=1e-5)
    assert int(logs["n_negatives"]) == int(g8[f"w8/rank{rank}/n_negatives"]) == 63
    assert float(logs["I_q_zm"]) == pytest.approx(float(g8[f"w8/rank{rank}/I_q_zm"]), rel=1e-5)
    d.mean().backward()
    checks = [(zr.grad, "g_z_hat"), (zpr.grad, "g_z_pos"), (ls.grad, "g_logit_scale")]
    if rank == 3:
        checks += [(proj["w1"].grad, "g_w1"), (proj["w2"].grad, "g_w2")]
    for got, want in checks:
        w = T(g8[f"w8/rank{rank}/{want}"])
        assert (got - w).abs().max() <= 1e-5 * w.abs().max() + 1e-9, want


def test_lossy_z_matches_reference(ref):
    g = ref["losses"]
    zh = T(g["lossy_z/z_hat"]).requires_grad_(True)
    d = olosses.lossy_z_distortion(zh, T(g["lossy_z/z"]))
    assert torch.allclose(d, T(g["lossy_z/distortion"]), rtol=1e-6)
    d.sum().backward()
    assert torch.equal(zh.grad, T(g["lossy_z/g_z_hat"]))


@pytest.mark.parametrize("mode", ["linear", "geometric", "constant"])
def test_annealer_matches_reference(ref, mode):
    """oracle Annealer and the package's Annealer against lossyless/helpers.py:768-870, both calling conventions."""
    from lossyless_b200.helpers import Annealer

    g = ref["losses"]
    a_o = olosses.Annealer(0.05 * 1e-5, 0.05, math.ceil(250 / 10), mode=mode)
    a_p = Annealer(0.05 * 1e-5, 0.05, n_steps_anneal=math.ceil(250 / 10), mode=mode)
    a_u = Annealer(0.05 * 1e-5, 0.05, n_steps_anneal=math.ceil(250 / 10), mode=mode)
    for i, s in enumerate(g["annealer/steps"]):
        want = float(g[f"annealer/{mode}"][i])
        assert a_o(int(s)) == pytest.approx(want, rel=1e-12)
        assert a_p(n_update_calls=int(s)) == pytest.approx(want, rel=1e-12)
        assert a_u(is_update=True) == pytest.approx(float(g[f"annealer_update/{mode}"][i]), rel=1e-12)


@pytest.mark.parametrize("distortion", ["lossy_Z", "contrastive", "direct"])
def test_rd_loss_matches_reference(ref, distortion):
    """LearnableCompressor.loss (learnable_compressors.py:241-275) at global_step 3 of a linear anneal: the value
    uses the final beta, the gradient the annealed one; oracle.rd_loss and the package's own .loss."""
    g = ref["step"]
    pre = f"{distortion}/loss_fn/"
    factor = 0.003 if distortion == "contrastive" else 1.0
    final_beta, labda = 0.05 * 1, 1 / factor
    beta = olosses.Annealer(final_beta * 1e-5, final_beta, math.ceil(100 / 10), mode="linear")(3)
    assert beta == pytest.approx(float(g[pre + "beta"]), rel=1e-12)
    rates, dists = T(g[pre + "rates"]).requires_grad_(True), T(g[pre + "dists"]).requires_grad_(True)
    loss, logs = olosses.rd_loss(rates, dists, beta, final_beta, labda)
    assert float(loss) == pytest.approx(float(g[pre + "loss"]), rel=1e-6)
    assert float(logs["loose_loss"]) == pytest.approx(float(g[pre + "loose_loss"]), rel=1e-6)
    loss.backward()
    assert torch.allclose(rates.grad, T(g[pre + "g_rates"]), rtol=1e-6)
    assert torch.allclose(dists.grad, T(g[pre + "g_dists"]), rtol=1e-6)

    # the package's LearnableCompressor.loss is host-side torch: same check without a GPU
    from lossyless_b200.learnable_compressors import LearnableCompressor

    class _M:
        final_beta_labda = (final_beta, labda)
        global_step = 3
        beta_annealer = staticmethod(lambda n_update_calls: beta)

    r2, d2 = T(g[pre + "rates"]).requires_grad_(True), T(g[pre + "dists"]).requires_grad_(True)
    loss_p, logs_p, _ = LearnableCompressor.loss(_M(), r2, d2)
    assert float(loss_p) == pytest.approx(float(g[pre + "loss"]), rel=1e-6)
    loss_p.backward()
    assert torch.allclose(r2.grad, T(g[pre + "g_rates"]), rtol=1e-6) and torch.allclose(d2.grad, T(g[pre + "g_dists"]), rtol=1e-6)


# ------------------------------------------------------------------------------------------ rates
@pytest.mark.parametrize("tag", TAGS)
def test_factorized_rate_matches_reference(ref, weights_dir, tag):
    """lossyless/rates.py HRateFactorizedPrior.{forward, compress, decompress, real_rate, aux_loss} on the published
    checkpoints: process_z_in/out, the per-row -log q sum, is_endToEnd=False's second (detached) pass, log keys."""
    g = ref["rates"]
    sd = torch.load(os.path.join(weights_dir, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)
    eb = oeb.split_state_dict(sd)
    b, s = sd["biasing"], sd["scaling"]
    z = T(g[f"{tag}/z"])
    # eval forward
    zh_in, q = oeb.forward(eb, oeb.process_z_in(z, b, s))
    assert torch.allclose(oeb.process_z_out(zh_in, b, s), T(g[f"{tag}/eval/z_hat"]), rtol=1e-6, atol=1e-7)
    rates = -torch.log(q).sum(-1)
    assert torch.allclose(rates, T(g[f"{tag}/eval/rates"]), rtol=1e-5)
    assert float(rates.mean() / math.log(2)) == pytest.approx(float(g[f"{tag}/eval/H_q_Z"]), rel=1e-5)
    # train forward: z_hat from the first noise draw, the rate from the second (detached) pass
    zh_in, _ = oeb.forward(eb, oeb.process_z_in(z, b, s), T(g[f"{tag}/train/noise0"]))
    assert torch.allclose(oeb.process_z_out(zh_in, b, s), T(g[f"{tag}/train/z_hat"]), rtol=1e-6, atol=1e-7)
    _, q = oeb.forward(eb, oeb.process_z_in(z, b, s), T(g[f"{tag}/train/noise1"]))
    assert torch.allclose(-torch.log(q).sum(-1), T(g[f"{tag}/train/rates"]), rtol=1e-5)
    # aux loss
    assert float(oeb.aux_loss(eb, sd["entropy_bottleneck.target"])) == pytest.approx(float(g[f"{tag}/aux_loss"]), rel=1e-6)
    # integer side: tables, bytes, decoded z, n_bits -- bit exact
    tab = oeb.update(eb)
    assert np.array_equal(tab["cdf"], g[f"{tag}/cdf"]) and np.array_equal(tab["cdf_length"], g[f"{tag}/cdf_length"])
    assert np.array_equal(tab["offset"], g[f"{tag}/offset"])
    strings, _ = oeb.compress(eb, tab, oeb.process_z_in(z, b, s))
    assert np.array_equal(np.array([len(x) for x in strings], dtype=np.int32), g[f"{tag}/lens"])
    assert b"".join(strings) == g[f"{tag}/bytes"].tobytes()
    z_dec, _ = oeb.decompress(eb, tab, strings)
    assert torch.equal(oeb.process_z_out(z_dec, b, s), T(g[f"{tag}/z_dec"]))
    assert 8 * np.mean([len(x) for x in strings]) == float(g[f"{tag}/n_bits"]) == float(g[f"{tag}/logs_n_bits"])


def test_hyperprior_rate_matches_reference(ref):
    """lossyless/rates.py:571-670 HRateHyperprior.forward_help in training mode (side MLP -> factorised bottleneck ->
    z MLP -> Gaussian conditional): values through the oracle's restatements."""
    g = ref["rates"]
    sd = {k[len("hyper/sd/"):]: T(g[k]) for k in g.files if k.startswith("hyper/sd/")}
    eb = oeb.split_state_dict(sd)

    def mlp(prefix, x):
        for i in (0, 4, 8):
            x = x @ sd[f"{prefix}.module.{i}.weight"].t() + sd[f"{prefix}.module.{i}.bias"]
            if i != 8:
                x = torch.relu(x)
        return x

    z = T(g["hyper/z"])
    z_in = oeb.process_z_in(z, sd["biasing"], sd["scaling"])
    side = mlp("side_encoder", z_in)
    side_hat, q_s = oeb.forward(eb, side, T(g["hyper/noise_side"]))
    scales, means = mlp("z_encoder", side_hat).chunk(2, -1)
    z_hat_in, q_z = og.forward(z_in, scales, means, T(g["hyper/noise_z"]))
    rates = -torch.log(q_s).sum(-1) - torch.log(q_z).sum(-1)
    assert torch.allclose(rates, T(g["hyper/rates"]), rtol=1e-5)
    assert torch.allclose(oeb.process_z_out(z_hat_in, sd["biasing"], sd["scaling"]), T(g["hyper/z_hat"]), rtol=1e-5, atol=1e-6)
    assert float(-torch.log(q_s).sum(-1).mean() / math.log(2)) == pytest.approx(float(g["hyper/H_q_S"]), rel=1e-5)
    # the reference's side bottleneck DOES receive a rate gradient (the q(s) term): the fixture must not be zero
    assert np.abs(g["hyper/g_eb_matrix0"]).max() > 0 and np.abs(g["hyper/g_side_w0"]).max() > 0


# ------------------------------------------------------------------------------------------ hub container
def test_container_matches_reference_writer(ref, weights_dir, native_lib):
    """hub/compressor.py compress_dataset wrote `container` from z_f16 (batches of 8, 20 images): the oracle codec fed
    with the same z gives the same file; the library's host-side container index reads it back."""
    g = ref["hub"]
    sd = torch.load(os.path.join(weights_dir, "beta5e-02_factorized_rate.pt"), map_location="cpu", weights_only=True)
    eb = oeb.split_state_dict(sd)
    tab = oeb.update(eb)
    z = T(g["z_f16"])
    strings, _ = oeb.compress(eb, tab, oeb.process_z_in(z, sd["biasing"], sd["scaling"]))
    blob = struct.pack(">I", 20) + b"".join(struct.pack(">I", len(s)) + s for s in strings)
    assert blob == g["container"].tobytes()
    zd, _ = oeb.decompress(eb, tab, strings)
    assert np.array_equal(oeb.process_z_out(zd, sd["biasing"], sd["scaling"]).numpy(), g["Z_hat"])
    assert float(g["rate5"]) == 8 * np.mean([len(s) for s in strings[:5]])
    assert np.array_equal(g["z_hat5"], g["Z_hat"][:5]) and np.array_equal(g["z_fwd5"], g["Z_hat"][:5])
    assert str(g["printed"]).startswith("Rate: %.2f bits/img | Encoding: " % (8 * len(blob) / 20))
    # labels: np.save(uint16), read back as int64
    import io

    assert np.array_equal(np.load(io.BytesIO(g["label_file"].tobytes())), (np.arange(20) % 7).astype(np.uint16))
    assert np.array_equal(g["Y"], np.arange(20) % 7) and g["Y"].dtype == np.int64
    # the library's container index on the reference's file
    from lossyless_b200 import _native as N

    raw = g["container"].copy()
    offsets, lens = np.zeros(21, dtype=np.int64), np.zeros(20, dtype=np.int32)
    N.call("ll_container_index", raw.ctypes.data, raw.size, 4, 20, offsets.ctypes.data, lens.ctypes.data)
    assert lens.tolist() == [len(s) for s in strings]
    assert all(raw[offsets[i]: offsets[i] + lens[i]].tobytes() == strings[i] for i in range(20))


# ------------------------------------------------------------------------------------------ downstream probe
@pytest.mark.parametrize("case", ["linear_clf", "mlp_clf", "mlp_reg"])
def test_predictor_matches_reference(golden_dir, case):
    """``lossyless_b200.predictors.Predictor`` (the consumer of the decompressed representations) against
    what the reference's own ``lossyless/predictors.py`` computed on the same parameters and inputs
    (tests/golden/ref_predictor.npz): train-mode ``step`` -- loss, every logged scalar, parameter gradients --, eval-mode
    ``step`` and the class probabilities of ``forward(is_logits=False)``."""
    import sys

    sys.path.insert(0, golden_dir)
    from ref_configs import predictor_hparams  # the hparams the fixture was made with

    from lossyless_b200.predictors import Predictor

    g = np.load(os.path.join(golden_dir, "ref_predictor.npz"))
    pre = case + "/"
    m = Predictor(predictor_hparams(case))
    sd = {k[len(pre + "sd/"):]: T(g[k]) for k in g.files if k.startswith(pre + "sd/")}
    assert set(sd) == set(m.state_dict()), "parameter names differ from the reference module's"
    m.load_state_dict(sd)
    x, y = T(g["x"]), T(g[pre + "y"])

    def close(a, b, tol=1e-5):
        return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= tol * max(1.0, abs(b))

    m.train()
    loss, logs = m.step((x, y))
    loss.backward()
    assert close(float(loss), float(g[pre + "train/loss"]))
    keys = [k[len(pre + "train/logs/"):] for k in g.files if k.startswith(pre + "train/logs/")]
    assert set(keys) == set(logs) - {"inference_time"}
    for k in keys:  # tasks_std of a single task is nan in the reference, and here
        assert close(float(logs[k]), float(g[pre + "train/logs/" + k])), k
    for n, p in m.named_parameters():
        want = T(g[pre + "train/grad/" + n])
        assert torch.allclose(p.grad, want, rtol=1e-4, atol=1e-7), n
    m.eval()
    with torch.no_grad():
        loss_e, logs_e = m.step((x, y))
        assert close(float(loss_e), float(g[pre + "eval/loss"]))
        for k in keys:
            assert close(float(logs_e[k]), float(g[pre + "eval/logs/" + k])), k
        assert torch.allclose(m(x, is_logits=False), T(g[pre + "eval/probs"]), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------ optimisers / schedulers
@pytest.mark.parametrize("sch,epochs", [("expdecay100", 6), ("unifmultistep100", 8)])
def test_configure_optimizers_matches_reference(golden_dir, sch, epochs):
    """``LearnableCompressor.configure_optimizers`` against the reference's own (tests/golden/ref_optim.npz, config/main.yaml's
    optimiser and scheduler groups): the same parameters in the same order under the main and the coder optimiser, the
    same hyper-parameters, the same scheduler classes and names, and -- stepping the schedulers once per epoch as ``fit``
    does at the end of every call -- the same learning rate after every epoch."""
    import sys

    sys.path.insert(0, golden_dir)
    from ref_configs import optim_hparams

    from lossyless_b200.learnable_compressors import LearnableCompressor

    g = np.load(os.path.join(golden_dir, "ref_optim.npz"))
    pre = sch + "/"
    m = LearnableCompressor(optim_hparams("contrastive", sch, epochs))
    names = {id(p): n for n, p in m.named_parameters()}
    optimizers, schedulers = m.configure_optimizers()
    assert len(optimizers) == int(g[pre + "n_optimizers"])
    assert [m.idcs_to_compressors[i] for i in range(len(optimizers))] == list(g[pre + "roles"])
    for i, opt in enumerate(optimizers):
        assert issubclass(type(opt), getattr(torch.optim, str(g[pre + f"opt{i}/class"])))
        assert [names[id(p)] for grp in opt.param_groups for p in grp["params"]] == list(g[pre + f"opt{i}/params"])
        grp = opt.param_groups[0]
        assert grp["lr"] == float(g[pre + f"opt{i}/lr"]) and grp["weight_decay"] == float(g[pre + f"opt{i}/weight_decay"])
        assert tuple(grp["betas"]) == tuple(g[pre + f"opt{i}/betas"]) and grp["eps"] == float(g[pre + f"opt{i}/eps"])
    assert [s["name"] for s in schedulers] == list(g[pre + "sched_names"])
    assert [type(s["scheduler"]).__name__ for s in schedulers] == list(g[pre + "sched_classes"])
    want = g[pre + "lr_after_epoch"]
    for e in range(epochs):
        for opt in optimizers:
            opt.step()
        for s in schedulers:
            s["scheduler"].step()
        got = [opt.param_groups[0]["lr"] for opt in optimizers]
        assert np.allclose(got, want[e], rtol=1e-12, atol=0), (e, got, want[e])
