"""Golden fixtures made by EXECUTING THE REFERENCE'S OWN PYTHON (/root/reference, in place) on seeded inputs.

Run here (the container that mounts /root/reference), from the repo root:

    python tests/golden/make_golden_ref.py

Writes tests/golden/ref_*.npz.  The reference's plain-torch files run verbatim; the two un-installable leaf
packages are stood in for by tests/golden/ref_shim.py (compressai arithmetic = the oracle's restatement;
clip.visual = the fp32 oracle ViT).  So these files pin, to reference code rather than to a restatement:

  ref_losses.npz  lossyless/distortions.py  ContrastiveDistortion.{compute_logits,compute_loss,forward} at W = 1
                  and W = 2 (two real gloo processes through lossyless/helpers.py GatherFromGpus, forward AND
                  backward), LossyZDistortion; lossyless/helpers.py Annealer (all modes);
                  lossyless/learnable_compressors.py LearnableCompressor.loss (value + gradient).
  ref_rates.npz   lossyless/rates.py HRateFactorizedPrior on the three published checkpoints (eval / train forward,
                  is_endToEnd=False second pass, compress / decompress / real_rate, aux_loss) and HRateHyperprior
                  (forward with gradients; its compress() defect at :691-692 is not exercised).
  ref_step.npz    LearnableCompressor.step end to end (MLP encoder; lossy_Z, contrastive and direct distortions):
                  parameters, inputs, noise draws, every logged scalar, gradients.
  ref_hub.npz     hub/compressor.py ClipCompressor.{compress, decompress, get_rate, compress_dataset,
                  decompress_dataset}: the container file bytes the reference's own writer produced.
  ref_predictor.npz  lossyless/predictors.py Predictor.{step, loss, forward} (python tests/golden/make_golden_ref.py predictor):
                  the downstream probe on 512-d representations, train (with gradients) and eval.
  ref_optim.npz   LearnableCompressor.configure_optimizers / get_specific_parameters + helpers.append_optimizer_scheduler_ /
                  get_lr_scheduler on config/main.yaml's optimiser and scheduler groups (... make_golden_ref.py optim):
                  the main / coder parameter split, optimiser hyper-parameters, learning rate after every epoch.
"""
import io
import math
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=FutureWarning)

import ref_shim  # noqa: E402
from ref_configs import CONTRASTIVE_KW, PREDICTOR_CASES, SeededImages, optim_hparams, predictor_hparams, step_hparams  # noqa: E402

TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]
WEIGHTS = os.path.join(ROOT, "lossyless_b200", "hub", "weights")


class AD(dict):
    """attribute dict, nested (the reference's NamespaceMap lives in utils/helpers.py behind Lightning imports)."""

    def __init__(self, d=()):
        super().__init__()
        for k, v in dict(d).items():
            self[k] = AD(v) if isinstance(v, dict) else v

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def npz(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()})
    print(name, os.path.getsize(path), "bytes,", len(d), "arrays")


def sd_np(prefix, module):
    return {f"{prefix}{k}": v for k, v in module.state_dict().items()}


# --------------------------------------------------------------------------------------------- losses
def _contrastive(rank_seed=None):
    from lossyless.distortions import ContrastiveDistortion

    torch.manual_seed(11)
    d = ContrastiveDistortion(**{**CONTRASTIVE_KW, "project_kwargs": dict(CONTRASTIVE_KW["project_kwargs"])})
    with torch.no_grad():  # a non-default temperature, above the 0.01 floor
        d.logit_scale.fill_(math.log(1 / 0.03))
    return d


def _w2_worker(rank, port, out, world=2, batch=8):
    """One real gloo process of the W = `world` contrastive run: the reference's own all_gather + backward."""
    import torch.distributed as dist

    warnings.filterwarnings("ignore", category=FutureWarning)
    ref_shim.import_reference()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d = _contrastive()
    g = torch.Generator().manual_seed(100 + rank)
    z_hat = torch.randn(batch, 512, generator=g).requires_grad_(True)
    z_pos = torch.randn(batch, 512, generator=g).requires_grad_(True)
    logits = d.compute_logits(z_hat, z_pos, None)
    dd, logs, _ = d(z_hat, z_pos, None, None)
    dd.mean().backward()
    res = dict(z_hat=z_hat.detach(), z_pos=z_pos.detach(), logits=logits.detach(), distortion=dd.detach(),
               I_q_zm=float(logs["I_q_zm"]), hat_H_m=float(logs["hat_H_m"]), n_negatives=int(logs["n_negatives"]),
               g_z_hat=z_hat.grad, g_z_pos=z_pos.grad, g_logit_scale=d.logit_scale.grad,
               g_w1=d.projector.module[0].weight.grad, g_w2=d.projector.module[-1].weight.grad)
    torch.save(res, out.format(rank))
    dist.barrier()
    dist.destroy_process_group()


def make_losses():
    import torch.multiprocessing as mp
    from lossyless.distortions import LossyZDistortion
    from lossyless.distributions import Deterministic
    from lossyless.helpers import Annealer

    out = {}
    # -- contrastive, one process (no process group: the gather branch is skipped, distortions.py:326)
    d = _contrastive()
    out.update(sd_np("contrastive/", d))
    g = torch.Generator().manual_seed(100)
    z_hat = torch.randn(8, 512, generator=g).requires_grad_(True)
    z_pos = torch.randn(8, 512, generator=g).requires_grad_(True)
    logits = d.compute_logits(z_hat, z_pos, None)
    h, logs = d.compute_loss(logits.clone())
    dd, logs2, _ = d(z_hat, z_pos, None, None)
    dd.mean().backward()
    out.update({"w1/z_hat": z_hat, "w1/z_pos": z_pos, "w1/logits": logits, "w1/hat_H_mlz": h, "w1/distortion": dd,
                "w1/I_q_zm": float(logs2["I_q_zm"]), "w1/hat_H_m": float(logs2["hat_H_m"]), "w1/n_negatives": int(logs2["n_negatives"]),
                "w1/g_z_hat": z_hat.grad, "w1/g_z_pos": z_pos.grad, "w1/g_logit_scale": d.logit_scale.grad,
                "w1/g_w1": d.projector.module[0].weight.grad, "w1/g_w2": d.projector.module[-1].weight.grad})
    # a second flavour: fixed temperature, no effective batch size, no projector, dot-product logits
    from lossyless.distortions import ContrastiveDistortion

    d2 = ContrastiveDistortion(temperature=0.5, is_train_temperature=False, is_cosine=False, effective_batch_size=None,
                               is_already_featurized=True, is_project=False)
    zz, zp = torch.randn(6, 32, generator=g), torch.randn(6, 32, generator=g)
    dd2, l2, _ = d2(zz, zp, None, None)
    out.update({"plain/z_hat": zz, "plain/z_pos": zp, "plain/distortion": dd2, "plain/I_q_zm": float(l2["I_q_zm"]),
                "plain/n_negatives": int(l2["n_negatives"])})

    # -- contrastive, two real gloo ranks
    with tempfile.TemporaryDirectory() as tmp:
        pat = os.path.join(tmp, "r{}.pt")
        port = 29000 + os.getpid() % 2000
        mp.spawn(_w2_worker, args=(port, pat), nprocs=2, join=True)
        for r in range(2):
            res = torch.load(pat.format(r), weights_only=False)
            out.update({f"w2/rank{r}/{k}": v for k, v in res.items()})

    # -- lossy Z (nn.PairwiseDistance(p=1) against p_Zlx.base_dist.mean)
    lz = LossyZDistortion(p_norm=1)
    z = torch.randn(8, 512, generator=g)
    zh = (z + 0.3 * torch.randn(8, 512, generator=g)).requires_grad_(True)
    p = Deterministic(z)
    dl, _, _ = lz(zh, None, p, None)
    dl.sum().backward()
    out.update({"lossy_z/z": z, "lossy_z/z_hat": zh, "lossy_z/distortion": dl, "lossy_z/g_z_hat": zh.grad})

    # -- Annealer, every mode, as LearnableCompressor.__init__ builds it (learnable_compressors.py:38-46)
    steps = np.arange(0, 40)
    for mode in ("linear", "geometric", "constant"):
        a = Annealer(0.05 * 1e-5, 0.05, n_steps_anneal=math.ceil(250 / 10), mode=mode)
        out[f"annealer/{mode}"] = np.array([a(n_update_calls=int(s)) for s in steps], dtype=np.float64)
        a2 = Annealer(0.05 * 1e-5, 0.05, n_steps_anneal=math.ceil(250 / 10), mode=mode)
        internal = []
        for _ in steps:  # the self-counting form (is_update=True)
            internal.append(a2(is_update=True))
        out[f"annealer_update/{mode}"] = np.array(internal, dtype=np.float64)
    out["annealer/steps"] = steps
    npz("ref_losses.npz", out)


# --------------------------------------------------------------------------------------------- rates
class _Parent:
    current_epoch = 0


def make_rates():
    from compressai.entropy_models import EntropyModel
    from lossyless.distributions import Deterministic
    from lossyless.rates import HRateFactorizedPrior, HRateHyperprior, get_rate_estimator

    out = {}
    for tag in TAGS:
        sd = torch.load(os.path.join(WEIGHTS, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)
        r = get_rate_estimator("H_factorized", z_dim=512, is_endToEnd=False,
                               kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))
        assert isinstance(r, HRateFactorizedPrior)
        r.load_state_dict(sd)
        g = torch.Generator().manual_seed(31)
        z = (torch.randn(12, 512, generator=g) * 0.45)
        z[10] *= 6.0  # escapes
        # eval forward (no coder tables yet -> no real rate in the logs)
        r.eval()
        with torch.no_grad():
            zh, rates, logs, _ = r(z, Deterministic(z), _Parent())
        out.update({f"{tag}/z": z, f"{tag}/eval/z_hat": zh, f"{tag}/eval/rates": rates, f"{tag}/eval/H_q_Z": float(logs["H_q_Z"])})
        # train forward: is_endToEnd=False => forward_help twice (second on detached z), two noise draws
        r.train()
        EntropyModel.noise_log.clear()
        torch.manual_seed(5)
        zg = z.clone().requires_grad_(True)
        zh, rates, logs, _ = r(zg, Deterministic(zg), _Parent())
        (rates.mean() + zh.sum() * 1e-3).backward()
        assert len(EntropyModel.noise_log) == 2
        out.update({f"{tag}/train/noise0": EntropyModel.noise_log[0][:, 0, :].t(), f"{tag}/train/noise1": EntropyModel.noise_log[1][:, 0, :].t(),
                    f"{tag}/train/z_hat": zh, f"{tag}/train/rates": rates, f"{tag}/train/H_q_Z": float(logs["H_q_Z"]),
                    f"{tag}/train/g_z": zg.grad, f"{tag}/train/g_scaling": r.scaling.grad, f"{tag}/train/g_biasing": r.biasing.grad,
                    f"{tag}/train/g_matrix0": r.entropy_bottleneck._matrix0.grad, f"{tag}/train/g_bias2": r.entropy_bottleneck._bias2.grad,
                    f"{tag}/train/g_factor1": r.entropy_bottleneck._factor1.grad})
        r.zero_grad()
        aux = r.aux_loss()
        aux.backward()
        out.update({f"{tag}/aux_loss": float(aux), f"{tag}/g_quantiles": r.entropy_bottleneck.quantiles.grad})
        # the real coder: prepare_compressor_ -> update(force) ; compress / decompress / real_rate
        r.eval()
        r.prepare_compressor_()
        with torch.no_grad():
            strings = r.compress(z)
            z_dec = r.decompress(strings)
            n_bits = r.real_rate(z)
            _, _, logs_e, _ = r(z, Deterministic(z), _Parent())
        assert len(strings) == 1 and "n_bits" in logs_e
        eb = r.entropy_bottleneck
        out.update({f"{tag}/cdf": eb._quantized_cdf, f"{tag}/cdf_length": eb._cdf_length, f"{tag}/offset": eb._offset,
                    f"{tag}/lens": np.array([len(s) for s in strings[0]], dtype=np.int32),
                    f"{tag}/bytes": np.frombuffer(b"".join(strings[0]), dtype=np.uint8),
                    f"{tag}/z_dec": z_dec, f"{tag}/n_bits": float(n_bits), f"{tag}/logs_n_bits": float(logs_e["n_bits"])})

    # -- H_hyper (lossyless/rates.py:571-670): seeded init, one training-mode forward with gradients
    torch.manual_seed(77)
    r = get_rate_estimator("H_hyper", z_dim=64, kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))
    assert isinstance(r, HRateHyperprior) and r.side_z_dim == 12  # z_dim 64 keeps the fixture small (MLPs of 256)
    g = torch.Generator().manual_seed(78)
    with torch.no_grad():
        r.scaling.add_(torch.randn(64, generator=g) * 0.1 + 0.5)
        r.biasing.add_(torch.randn(64, generator=g) * 0.1)
    out.update(sd_np("hyper/sd/", r))
    z = torch.randn(8, 64, generator=g) * 0.45
    r.train()
    EntropyModel.noise_log.clear()
    torch.manual_seed(6)
    zg = z.clone().requires_grad_(True)
    zh, rates, logs, _ = r(zg, Deterministic(zg), _Parent())
    (rates.mean() + zh.sum() * 1e-3).backward()
    assert len(EntropyModel.noise_log) == 2  # side bottleneck [C,1,B], then the Gaussian conditional [B,C]
    out.update({"hyper/z": z, "hyper/noise_side": EntropyModel.noise_log[0][:, 0, :].t(), "hyper/noise_z": EntropyModel.noise_log[1],
                "hyper/z_hat": zh, "hyper/rates": rates, "hyper/H_q_Z": float(logs["H_q_Z"]), "hyper/H_q_S": float(logs["H_q_S"]),
                "hyper/H_q_ZlS": float(logs["H_q_ZlS"]), "hyper/g_z": zg.grad, "hyper/g_scaling": r.scaling.grad,
                "hyper/g_eb_matrix0": r.entropy_bottleneck._matrix0.grad, "hyper/g_eb_bias4": r.entropy_bottleneck._bias4.grad,
                "hyper/g_eb_factor0": r.entropy_bottleneck._factor0.grad,
                "hyper/g_side_w0": r.side_encoder.module[0].weight.grad, "hyper/g_zenc_wlast": r.z_encoder.module[-1].weight.grad})
    r.eval()
    with torch.no_grad():
        zh, rates, logs, _ = r(z, Deterministic(z), _Parent())
    out.update({"hyper/eval/z_hat": zh, "hyper/eval/rates": rates})
    npz("ref_rates.npz", out)


# --------------------------------------------------------------------------------------------- step
def make_step():
    from compressai.entropy_models import EntropyModel
    from lossyless.learnable_compressors import LearnableCompressor

    out = {}
    for distortion in ("lossy_Z", "contrastive", "direct"):
        hp = step_hparams(distortion)
        torch.manual_seed(3)
        m = LearnableCompressor(AD(hp))
        g = torch.Generator().manual_seed(4)
        with torch.no_grad():
            m.rate_estimator.scaling.add_(torch.randn(16, generator=g) * 0.1 + 0.7)
            m.rate_estimator.biasing.add_(torch.randn(16, generator=g) * 0.1)
        m.global_step = 3
        m.train()
        x = torch.randn(10, 24, generator=g)
        x_pos = x + 0.1 * torch.randn(10, 24, generator=g)
        aux = x_pos if distortion == "contrastive" else x
        EntropyModel.noise_log.clear()
        torch.manual_seed(9)
        loss, logs, other = m.step((x, (None, aux)))
        loss.backward()
        pre = f"{distortion}/"
        out.update({pre + "sd/" + k: v for k, v in m.state_dict().items()})
        out.update({pre + "x": x, pre + "aux": aux, pre + "loss": float(loss), pre + "z": other["z"],
                    pre + "n_noise": len(EntropyModel.noise_log)})
        for i, nz in enumerate(EntropyModel.noise_log):
            out[pre + f"noise{i}"] = nz[:, 0, :].t()
        for k, v in logs.items():
            out[pre + "logs/" + k] = float(v)
        for n, p in m.named_parameters():
            if p.grad is not None:
                out[pre + "grad/" + n] = p.grad
        # the loss in isolation (value uses final beta, gradient uses the annealed beta)
        rates = torch.rand(10, generator=g).mul(50).requires_grad_(True)
        dists = torch.rand(10, generator=g).requires_grad_(True)
        l2, lg2, _ = m.loss(rates, dists)
        l2.backward()
        out.update({pre + "loss_fn/rates": rates, pre + "loss_fn/dists": dists, pre + "loss_fn/loss": float(l2), pre + "loss_fn/g_rates": rates.grad,
                    pre + "loss_fn/g_dists": dists.grad, pre + "loss_fn/beta": float(lg2["beta"]), pre + "loss_fn/loose_loss": float(lg2["loose_loss"])})
        # eval-mode step + the test-time coder (real rate keys)
        m.eval()
        m.on_test_epoch_start()
        with torch.no_grad():
            loss_e, logs_e, _ = m.step((x, (None, aux)))
        out.update({pre + "eval/loss": float(loss_e), pre + "eval/n_bits": float(logs_e["n_bits"]), pre + "eval/rate": float(logs_e["rate"]),
                    pre + "eval/distortion": float(logs_e["distortion"]), pre + "eval/keys": np.array(sorted(logs_e.keys()))})
    npz("ref_step.npz", out)


# --------------------------------------------------------------------------------------------- hub
def make_hub():
    _, hub = ref_shim.import_reference()
    out = {}
    tag = "beta5e-02"
    sd = torch.load(os.path.join(WEIGHTS, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)
    # device "cpu:0": the reference refuses the literal "cpu" (hub/compressor.py:180) only because its CLIP is fp16
    comp = hub.ClipCompressor(dict(sd), device="cpu:0")
    ds = SeededImages(20)
    with tempfile.TemporaryDirectory() as tmp:
        f, lf = os.path.join(tmp, "Z.bin"), os.path.join(tmp, "Y.npy")
        buf = io.StringIO()
        so, sys.stdout = sys.stdout, buf
        try:
            comp.compress_dataset(ds, f, label_file=lf, kwargs_dataloader=dict(batch_size=8, num_workers=0), is_info=True)
            Z_hat, Y = comp.decompress_dataset(f, label_file=lf, is_info=True, is_cpu=False)
        finally:
            sys.stdout = so
        out["container"] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
        out["label_file"] = np.frombuffer(open(lf, "rb").read(), dtype=np.uint8)
        out["printed"] = np.array(buf.getvalue())
    X = torch.stack([ds[i][0] for i in range(20)]).half()
    with torch.no_grad():
        z = comp.clip(X)
        strings = comp.compress(X[:5])
        rate = comp.get_rate(X[:5])
        z_hat5 = comp.decompress(strings)
        z_fwd = comp(X[:5])
    out.update(dict(z_f16=z, Z_hat=Z_hat, Y=Y, rate5=float(rate), z_hat5=z_hat5, z_fwd5=z_fwd,
                    lens5=np.array([len(s) for s in strings], dtype=np.int32)))
    npz("ref_hub.npz", out)


def make_losses_w8():
    """SURVEY 8(d) configuration 3 trains on 8 ranks: the reference's ContrastiveDistortion over EIGHT real gloo processes
    (4 samples per rank: logits [8, 64] per rank, n_negatives = 63).  Its own small fixture (ref_losses_w8.npz) so that
    ref_losses.npz stays byte for byte what earlier rounds committed."""
    import torch.multiprocessing as mp

    out = {}  # the module's parameters are ref_losses.npz's "contrastive/*" (same seed)
    with tempfile.TemporaryDirectory() as tmp:
        pat = os.path.join(tmp, "r{}.pt")
        port = 31000 + os.getpid() % 2000
        mp.spawn(_w2_worker, args=(port, pat, 8, 4), nprocs=8, join=True)
        for r in range(8):
            res = torch.load(pat.format(r), weights_only=False)
            # every rank's inputs, three ranks' results, one rank's (large) projector gradients
            keep = [k for k in res if r in (0, 3, 7) and (r == 3 or k not in ("g_w1", "g_w2"))] if r in (0, 3, 7) else ["z_hat", "z_pos"]
            out.update({f"w8/rank{r}/{k}": res[k] for k in keep})
    npz("ref_losses_w8.npz", out)


# --------------------------------------------------------------------------------------------- predictor
def make_predictor():
    """lossyless/predictors.py Predictor (the downstream consumer of Z_hat): ``step`` in train mode (loss, every
    logged scalar, gradients) and in eval mode, ``forward`` as probabilities, for a linear and an MLP classification probe
    on 512-d representations and an MLP regression probe.  (torchmetrics' ``accuracy`` is the shim's two-line restatement.)"""
    from lossyless.predictors import Predictor

    out = {}
    for case, c in PREDICTOR_CASES.items():
        torch.manual_seed(31)
        m = Predictor(AD(predictor_hparams(case)))
        g = torch.Generator().manual_seed(32)
        x = torch.randn(48, 512, generator=g)
        y = torch.randint(0, 10, (48,), generator=g) if c["is_clf"] else torch.randn(48, 3, generator=g)
        pre = case + "/"
        out.update({pre + "sd/" + k: v.clone() for k, v in m.state_dict().items()})
        out["x"] = x  # the same seeded representations for every case
        out[pre + "y"] = y
        m.train()
        loss, logs = m.step((x, y))
        loss.backward()
        out[pre + "train/loss"] = float(loss)
        for k, v in logs.items():
            if k != "inference_time":
                out[pre + "train/logs/" + k] = float(v)
        for n, p in m.named_parameters():
            out[pre + "train/grad/" + n] = p.grad.clone()
        m.eval()
        with torch.no_grad():
            loss_e, logs_e = m.step((x, y))
            out[pre + "eval/loss"] = float(loss_e)
            for k, v in logs_e.items():
                if k != "inference_time":
                    out[pre + "eval/logs/" + k] = float(v)
            out[pre + "eval/probs"] = m(x, is_logits=False)
    npz("ref_predictor.npz", out)


# --------------------------------------------------------------------------------------------- optimisers / schedulers
def make_optim():
    """lossyless/learnable_compressors.py ``configure_optimizers`` / ``get_specific_parameters`` with lossyless/helpers.py
    ``append_optimizer_scheduler_ / get_optimizer / get_lr_scheduler`` on config/main.yaml's groups: which parameters each
    optimiser owns (main vs coder), its hyper-parameters, and the learning rate after every epoch (``scheduler.step()``
    once per epoch, Lightning's default interval)."""
    from lossyless.learnable_compressors import LearnableCompressor

    out = {}
    for sch, epochs in (("expdecay100", 6), ("unifmultistep100", 8)):
        torch.manual_seed(3)
        m = LearnableCompressor(AD(optim_hparams("contrastive", sch, epochs)))
        names = {id(p): n for n, p in m.named_parameters()}
        optimizers, schedulers = m.configure_optimizers()
        pre = sch + "/"
        out[pre + "n_optimizers"] = len(optimizers)
        out[pre + "roles"] = np.array([m.idcs_to_compressors[i] for i in range(len(optimizers))])
        for i, opt in enumerate(optimizers):
            out[pre + f"opt{i}/class"] = np.array(type(opt).__name__)
            out[pre + f"opt{i}/params"] = np.array([names[id(p)] for g_ in opt.param_groups for p in g_["params"]])
            out[pre + f"opt{i}/lr"] = opt.param_groups[0]["lr"]
            out[pre + f"opt{i}/weight_decay"] = opt.param_groups[0]["weight_decay"]
            out[pre + f"opt{i}/betas"] = np.array(opt.param_groups[0]["betas"])
            out[pre + f"opt{i}/eps"] = opt.param_groups[0]["eps"]
        out[pre + "sched_names"] = np.array([s_["name"] for s_ in schedulers])
        out[pre + "sched_classes"] = np.array([type(s_["scheduler"]).__name__ for s_ in schedulers])
        lrs = []
        for _ in range(epochs):
            for opt in optimizers:
                opt.step()  # no gradients: nothing moves; keeps torch's "scheduler before optimizer" check quiet
            for s_ in schedulers:
                s_["scheduler"].step()
            lrs.append([opt.param_groups[0]["lr"] for opt in optimizers])
        out[pre + "lr_after_epoch"] = np.array(lrs, dtype=np.float64)
    npz("ref_optim.npz", out)


if __name__ == "__main__":
    ref_shim.import_reference()
    which = sys.argv[1:] or ["losses", "rates", "step", "hub"]
    for w in which:
        globals()[f"make_{w}"]()
