"""``LearnableCompressor``: the rate-distortion model surface of the reference
(lossyless/learnable_compressors.py:23-450) without PyTorch-Lightning (not installable here):
same constructor argument (an ``hparams`` attribute dict), ``forward / step / loss /
final_beta_labda / training_step / validation_step / test_step / configure_optimizers /
get_specific_parameters / on_test_epoch_start / set_featurize_mode_``, same log keys.

``fit`` is the thin replacement of ``Trainer.fit`` for this path: per batch one optimiser pass per
entry of ``idcs_to_compressors`` ("main", then "coder"), gradient clipping, optional data-parallel
gradient all-reduce (buckets only -- the encode path itself has no collective).
"""
import math
import os

import torch
import torch.nn as nn

from .architectures import get_Architecture
from .distortions import get_distortion_estimator
from .distributions import CondDist
from .helpers import BASE_LOG, Annealer, NamespaceMap, OrderedSet, Timer, append_optimizer_scheduler_
from .rates import get_rate_estimator

__all__ = ["LearnableCompressor", "default_clip_hparams", "allreduce_mean_"]


class LearnableCompressor(nn.Module):
    """Main network for learning a neural compression."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams if isinstance(hparams, NamespaceMap) else NamespaceMap(hparams)
        self.global_step, self.current_epoch = 0, 0
        self.logged = {}

        self.p_ZlX = self.get_encoder()
        self.rate_estimator = self.get_rate_estimator()
        self.distortion_estimator = self.get_distortion_estimator()
        if self.hparams.evaluation.featurizer.is_online:
            raise NotImplementedError("the online evaluator (lossyless/predictors.py) is outside the hot path")

        #! careful this is actually 1/beta from the paper
        final_beta = self.final_beta_labda[0]
        self.beta_annealer = Annealer(final_beta * 1e-5, final_beta,
                                      n_steps_anneal=math.ceil(1 / 10 * self.hparams.data.max_steps),
                                      mode=self.hparams.featurizer.loss.beta_anneal)
        self.is_features = self.hparams.featurizer.is_features
        self.out_shape = (self.hparams.encoder.z_dim,) if self.is_features else self.hparams.data.shape
        self.stage = self.hparams.stage

    # -- construction -------------------------------------------------------------------
    def get_encoder(self):
        cfg = self.hparams.encoder
        return CondDist(self.hparams.data.shape, cfg.z_dim, family=cfg.fam,
                        Architecture=get_Architecture(cfg.arch, **cfg.arch_kwargs), **cfg.fam_kwargs)

    def get_rate_estimator(self):
        cfg = self.hparams.rate
        rate_estimator = get_rate_estimator(cfg.mode, p_ZlX=self.p_ZlX, **cfg.kwargs)
        rate_estimator.make_pickable_()  # pickable before any process spawn
        return rate_estimator

    def get_distortion_estimator(self):
        cfg = self.hparams.distortion
        return get_distortion_estimator(cfg.mode, **cfg.kwargs)

    # -- representation -----------------------------------------------------------------
    def forward(self, x, is_compress=False, is_features=None, is_dist=False):
        """Represent ``x``: codes / (de)quantised features, or the reconstruction."""
        if is_features is None:
            is_features = self.is_features
        cache = getattr(self, "_view_cache", None)
        p_Zlx = cache[1] if cache is not None and cache[0] is x else self.p_ZlX(x)
        z = p_Zlx.rsample()
        if is_compress:
            z_hat = self.rate_estimator.compress(z, self)
        else:
            z_hat, rates, r_logs, r_other = self.rate_estimator(z, p_Zlx, self)
            if is_dist and "z_to_dist" in r_other:
                z_hat = r_other["z_to_dist"]
        if is_features:
            return z_hat
        return self.distortion_estimator.q_YlZ(z_hat)  # direct distortion only

    def step(self, batch, is_rate_only=False):
        x, targets = batch
        self._view_cache = None
        # encoder_time is only reported next to the real coder's n_bits (evaluation): a training step keeps the host
        # running ahead of the device instead of draining it on both sides of the encoder
        with Timer(sync=not self.training) as encoder_timer:
            x_pos = targets[1] if isinstance(targets, (tuple, list)) and len(targets) > 1 else None
            if not is_rate_only and self._can_batch_views(x, x_pos):
                # the contrastive distortion encodes the positive view through `parent(x_pos)` later in this step
                # (lossyless/distortions.py:305-318): both views go through the CLIP tower in ONE call (twice the
                # GEMM rows per launch, half the weight-gradient GEMMs when it is unfrozen); forward() below hands
                # the second half back when it is asked for x_pos.  Same values: the tower has no batch statistics.
                p_Zlx, p_pos = self.p_ZlX.forward_many([x, x_pos])
                self._view_cache = (x_pos, p_pos)
            else:
                p_Zlx = self.p_ZlX(x)  # batch shape [B], event shape [z_dim]
        z = p_Zlx.rsample()
        z_hat, rates, r_logs, r_other = self.rate_estimator(z, p_Zlx, self)

        if "n_bits" in r_logs:
            batch_size = x.size(0)
            r_logs["encoder_time"] = encoder_timer.duration / batch_size
            r_logs["sender_time"] = r_logs["encoder_time"] + r_logs["compress_time"]
            if self.hparams.data.mode == "image":
                _, __, height, width = x.shape
                r_logs["bpp"] = r_logs["n_bits"] / (height * width)

        if is_rate_only:
            r_logs["rate"] = rates.mean() / math.log(BASE_LOG)
            return rates, r_logs, r_other

        z_to_dist = r_other.pop("z_to_dist") if "z_to_dist" in r_other else z_hat
        _, aux_target = targets
        distortions, d_logs, d_other = self.distortion_estimator(z_to_dist, aux_target, p_Zlx, self)
        loss, logs, other = self.loss(rates, distortions)

        logs.update(r_logs)
        logs.update(d_logs)
        logs.update(dict(zmin=z_hat.min(), zmax=z_hat.max(), zmean=z_hat.mean()))
        other.update(r_other)
        other.update(d_other)
        # the reference copies this sample to the host here (for its plotting callbacks); a .cpu() per step is a full
        # host-device synchronisation, so the tensor stays on the device and consumers copy it if they want it
        other["X"] = x[0].detach()
        other.update(dict(z=z.detach()))
        self._view_cache = None
        return loss, logs, other

    def _can_batch_views(self, x, x_pos):
        from .architectures import PretrainedSSL
        from .distortions import ContrastiveDistortion

        de = self.distortion_estimator
        return (isinstance(de, ContrastiveDistortion) and not de.is_already_featurized and torch.is_tensor(x_pos)
                and isinstance(self.p_ZlX.mapper, PretrainedSSL) and x_pos.shape == x.shape and x_pos.dtype == x.dtype)

    @property
    def final_beta_labda(self):
        cfg = self.hparams
        labda = 1 / cfg.distortion.factor_beta
        beta = cfg.featurizer.loss.beta * cfg.rate.factor_beta
        return beta, labda

    def loss(self, rates, distortions):
        rates, distortions = rates.float(), distortions.float()
        curr_beta = self.beta_annealer(n_update_calls=self.global_step)
        final_beta, labda = self.final_beta_labda
        loose_loss = (labda * distortions + final_beta * rates).mean().detach()
        rate, distortion = rates.mean(0), distortions.mean(0)
        # annealed beta drives the gradient, final beta the reported value
        beta_rate = curr_beta * rate
        beta_rate = beta_rate - beta_rate.detach() + (final_beta * rate.detach())
        loss = labda * distortion + beta_rate
        logs = dict(loose_loss=loose_loss / math.log(BASE_LOG), loss=loss / math.log(BASE_LOG),
                    rate=rate / math.log(BASE_LOG), distortion=distortion / math.log(BASE_LOG), beta=curr_beta)
        logs["ratedist"] = logs["rate"] + logs["distortion"]
        return loss, logs, dict()

    # -- Lightning-style hooks -----------------------------------------------------------
    def log_dict(self, d, **kw):
        if getattr(self, "_defer_log_sync", False):
            # inside fit(): keep the 0-dim tensors; fit() turns them into floats once, at its end -- a float() per
            # logged scalar per step makes the host wait for the GPU between the forward and the backward pass
            self.logged.update({k: (v.detach() if torch.is_tensor(v) else v) for k, v in d.items()})
        else:
            self.logged.update({k: (float(v) if torch.is_tensor(v) else v) for k, v in d.items()})

    def training_step(self, batch, batch_idx, optimizer_idx=0):
        curr_step = self.idcs_to_compressors[optimizer_idx]
        if curr_step == "main":
            loss, logs, other = self.step(batch)
            self._save = other
        elif curr_step == "coder":
            loss = self.rate_estimator.aux_loss()
            logs = dict(coder_loss=loss)
        else:
            raise ValueError(f"Unkown curr_step={curr_step}.")
        self.log_dict({f"train/{self.stage}/{k}": v for k, v in logs.items()}, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx):
        loss, logs, _ = self.step(batch)
        self.log_dict({f"val/{self.stage}/{k}": v for k, v in logs.items()}, on_epoch=True, on_step=False, sync_dist=True)
        return loss

    def test_step(self, batch, batch_idx):
        is_communicate = self.stage == "comm"
        loss, logs, _ = self.step(batch, is_rate_only=is_communicate)
        self.log_dict({f"test/{self.stage}/{k}": v for k, v in logs.items()}, on_epoch=True, on_step=False, sync_dist=True)
        return loss

    def on_test_epoch_start(self):
        self.rate_estimator.prepare_compressor_()

    def get_specific_parameters(self, mode):
        all_param = OrderedSet(self.parameters())
        coder_param = OrderedSet(self.rate_estimator.aux_parameters())
        rate_param = OrderedSet(self.rate_estimator.parameters()) - coder_param
        online_param = OrderedSet([])
        notrate_param = all_param - (coder_param | online_param | rate_param)
        try:
            return dict(all=all_param, coder=coder_param, rate=rate_param, online=online_param,
                        main=notrate_param | rate_param, notrate=notrate_param)[mode]
        except KeyError:
            raise ValueError(f"Unkown parameter mode={mode}.") from None

    def configure_optimizers(self):
        self.idcs_to_compressors = {}
        optimizers, schedulers = [], []
        self.idcs_to_compressors[0] = "main"
        append_optimizer_scheduler_(self.hparams.optimizer_feat, self.hparams.get("scheduler_feat"),
                                    self.get_specific_parameters("main"), optimizers, schedulers, name="lr_featurizer")
        coder_parameters = self.get_specific_parameters("coder")
        if len(coder_parameters) > 0:
            self.idcs_to_compressors[1] = "coder"
            append_optimizer_scheduler_(self.hparams.optimizer_coder, self.hparams.get("scheduler_coder"),
                                        coder_parameters, optimizers, schedulers, name="lr_coder")
        return optimizers, schedulers

    def set_featurize_mode_(self):
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()
        self.rate_estimator.make_pickable_()

    # -- minimal trainer -----------------------------------------------------------------
    # -- checkpoints (Lightning's file layout, reference main.py:516-547) ------------------
    def save_checkpoint(self, path, weights_only=True):
        """``trainer.save_checkpoint(path, weights_only=True)``: ``state_dict`` + ``hyper_parameters`` + step counters
        under Lightning's key names.  ``hparams_name`` is written as Lightning 1.x writes it for
        ``save_hyperparameters(hparams)`` (``hparams_type`` only exists for omegaconf containers): without ``hparams_name`` its ``_load_model_state`` would spread the
        hyper-parameters as keyword arguments instead of passing the one ``hparams`` container the reference's
        ``LearnableCompressor.__init__(hparams)`` takes.  Non-plain entries are made portable: ``torch.dtype`` values
        become strings, an embedded CLIP state dict (``encoder.arch_kwargs.clip_state_dict``) is dropped -- the
        weights are in ``state_dict`` already."""
        self.rate_estimator.make_pickable_()
        ckpt = dict(state_dict=self.state_dict(), hyper_parameters=_plain(self.hparams), global_step=self.global_step,
                    epoch=self.current_epoch, hparams_name="hparams")
        ckpt["pytorch-lightning_version"] = "1.3.8"  # what requirements/environment.yaml pins; informational
        torch.save(ckpt, path)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location="cpu", strict=True, **kwargs):
        """``LightningModule.load_from_checkpoint``: rebuild from ``hyper_parameters`` (``kwargs`` override
        top-level entries, as Lightning's do) and load ``state_dict``.  Accepts a path or an already loaded
        checkpoint dict."""
        ckpt = checkpoint_path if isinstance(checkpoint_path, dict) else torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        if "state_dict" not in ckpt or "hyper_parameters" not in ckpt:
            raise KeyError("not a Lightning checkpoint: need 'state_dict' and 'hyper_parameters'")
        hp = _unplain(dict(ckpt["hyper_parameters"]))
        hp.update(kwargs)
        model = cls(NamespaceMap(hp))
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        model.global_step = int(ckpt.get("global_step", 0))
        model.current_epoch = int(ckpt.get("epoch", 0))
        return model

    def fit(self, batches, max_steps=None, grad_clip=3.0, process_group=None, grad_comm_dtype=None, callbacks=(),
            skip_grad_sync=False):
        """Lightning-free ``Trainer.fit``, ONE EPOCH per call: per batch, one pass per optimiser ("main" then
        "coder"); ``gradient_clip_val=3`` as config/main.yaml; with ``torch.distributed`` initialised the
        trainable gradients are all-reduced (mean) in flat buckets per pass (``allreduce_mean_``;
        ``grad_comm_dtype=torch.bfloat16`` for the bf16 buckets of an unfrozen encoder).

        Optimisers and schedulers are built ONCE (first call; ``configure_optimizers``) and kept on the module, so
        AdamW moments survive across calls; at the end of every call the schedulers step (Lightning's default
        ``interval="epoch"``) and ``current_epoch`` advances -- ``expdecay`` / ``MultiStepLR`` schedules and
        ``warmup_k_epoch`` see the epochs go by.  ``reset_optimizers_()`` drops the cached state (e.g. after
        un-freezing parameters)."""
        state = getattr(self, "_trainer_state", None)
        if state is None:
            for cb in callbacks:  # BaseFinetuning hooks run before the optimisers see the parameters
                if hasattr(cb, "freeze_before_training"):
                    cb.freeze_before_training(self)
            optimizers, schedulers = self.configure_optimizers()
            state = self._trainer_state = dict(optimizers=optimizers, schedulers=schedulers)
        optimizers, schedulers = state["optimizers"], state["schedulers"]
        # skip_grad_sync: timing A/B only (bench.py --mode train measures the exposed collective time with it)
        is_dist = torch.distributed.is_available() and torch.distributed.is_initialized() and not skip_grad_sync
        self.train()
        # no host synchronisation inside the loop (see log_dict); LL_SYNC_LOGS=1 restores a float() per logged scalar (A/B)
        self._defer_log_sync = os.environ.get("LL_SYNC_LOGS", "0") != "1"
        self._loss_steps = 0
        try:
            self._fit_loop(batches, optimizers, max_steps, grad_clip, is_dist, process_group, grad_comm_dtype)
        finally:
            self._defer_log_sync = False
            self.logged = {k: (float(v) if torch.is_tensor(v) else v) for k, v in self.logged.items()}
            if getattr(self, "_loss_host", None) is not None:
                if torch.cuda.is_available():
                    torch.cuda.current_stream().synchronize()
                self.loss_history = self._loss_host[: getattr(self, "_loss_steps", 0)].clone()
        for s in schedulers:
            s["scheduler"].step()
        self.current_epoch += 1
        return self.logged

    def reset_optimizers_(self):
        """Forget the optimisers / schedulers ``fit`` keeps between calls (they are rebuilt on the next call)."""
        self._trainer_state = None

    def _fit_loop(self, batches, optimizers, max_steps, grad_clip, is_dist, process_group, grad_comm_dtype):
        from .helpers import FlatAdamW

        world = torch.distributed.get_world_size(process_group) if is_dist else 1
        for batch_idx, batch in enumerate(_prefetch_to_device(batches, getattr(next(self.parameters(), None), "device", "cpu"))):
            if max_steps is not None and self.global_step >= max_steps:
                break
            for oi, opt in enumerate(optimizers):
                flat = isinstance(opt, FlatAdamW)
                opt.zero_grad(set_to_none=True)
                if flat:
                    # an unfrozen CLIP tower: its gradient slices are all-reduced from INSIDE the backward pass, block
                    # by block, under the GEMMs of the remaining blocks
                    opt.prepare_backward(sync=is_dist, group=process_group, world=world)
                loss = self.training_step(batch, batch_idx, optimizer_idx=oi)
                if oi == 0 and loss.is_cuda:
                    # the step's loss goes to the host every step -- as an asynchronous copy into pinned memory
                    # (``loss_history``, valid once fit() returns), not as a float() that would stall the host
                    hist = getattr(self, "_loss_host", None)
                    if hist is None or hist.numel() <= batch_idx:
                        new = torch.zeros(max(1024, 2 * (batch_idx + 1)), dtype=torch.float32).pin_memory()
                        if hist is not None:
                            torch.cuda.current_stream().synchronize()
                            new[: hist.numel()] = hist
                        hist = self._loss_host = new
                    hist[batch_idx: batch_idx + 1].copy_(loss.detach().reshape(1), non_blocking=True)
                    self._loss_steps = batch_idx + 1
                loss.backward()
                params = [p for g in opt.param_groups for p in g["params"] if p.grad is not None]
                if is_dist and params:
                    # (the tower's own gradients are 16-bit already; the few small float32 ones left here stay float32)
                    allreduce_mean_(params, process_group, comm_dtype=None if flat else grad_comm_dtype)
                if flat:
                    opt.finish_backward()
                    if grad_clip:
                        opt.clip_grad_norm_(grad_clip)
                elif grad_clip:
                    torch.nn.utils.clip_grad_norm_(params, grad_clip)
                opt.step()
            self.global_step += 1


def _map_tensors(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_tensors(o, fn) for o in obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    return obj


def _prefetch_to_device(batches, device):
    """``fit``'s input side: batches whose tensors are already on ``device`` pass through untouched; HOST batches (what a
    DataLoader hands over -- the reference lets Lightning copy each one synchronously at the top of its step) are
    copied on a side stream ONE BATCH AHEAD, so the copy of batch i + 1 runs under the kernels of batch i (308 MB of
    float32 per 2 x 256 images is 5.6 ms of link time against a 21 ms step).  Pinned host memory makes the copy
    asynchronous; pageable memory still works (the copy then blocks the host, as in the reference)."""
    device = torch.device(device)
    if device.type != "cuda":
        yield from batches
        return
    side = None
    it = iter(batches)

    def is_host(b):
        found = []
        _map_tensors(b, lambda t: found.append(t.device.type == "cpu") or t)
        return any(found)

    def stage(b):
        nonlocal side
        if not is_host(b):
            return b, None
        if side is None:
            side = torch.cuda.Stream(device)
        main = torch.cuda.current_stream(device)
        with torch.cuda.stream(side):
            moved = _map_tensors(b, lambda t: t.to(device, non_blocking=True))
            ready = torch.cuda.Event()
            ready.record(side)
        _map_tensors(moved, lambda t: t.record_stream(main) or t)
        return moved, ready

    try:
        nxt = stage(next(it))
    except StopIteration:
        return
    while nxt is not None:
        cur, ready = nxt
        try:
            nxt = stage(next(it))  # batch i + 1 starts moving before batch i's kernels are enqueued
        except StopIteration:
            nxt = None
        if ready is not None:
            torch.cuda.current_stream(device).wait_event(ready)
        yield cur


def _plain(d, key=None):
    """NamespaceMap -> nested plain dicts (what ends up under 'hyper_parameters'): ``torch.dtype`` values become
    strings ("torch.bfloat16"), embedded state dicts (``clip_state_dict``) are dropped."""
    if key == "clip_state_dict":
        return None
    if isinstance(d, dict):
        return {k: _plain(v, k) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return type(d)(_plain(v) for v in d)
    if isinstance(d, torch.dtype):
        return str(d)
    return d


def _unplain(d):
    """Inverse of ``_plain`` for the dtype strings."""
    if isinstance(d, dict):
        return {k: _unplain(v) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return type(d)(_unplain(v) for v in d)
    if isinstance(d, str) and d.startswith("torch.") and isinstance(getattr(torch, d[6:], None), torch.dtype):
        return getattr(torch, d[6:])
    return d


def allreduce_mean_(params, group=None, comm_dtype=None, bucket_bytes=32 << 20):
    """DDP's job for this path: the trainable gradients averaged over ranks, the only collective of the
    training step besides the logits-tile gather.  Gradients are packed in parameter order into flat
    buckets of at most ``bucket_bytes`` (one bucket for the frozen-encoder recipe's ~131 k parameters; a few
    for the 88 M of an unfrozen CLIP tower), every bucket's NCCL all-reduce is launched asynchronously
    before the first one is waited on, and ``comm_dtype`` (e.g. ``torch.bfloat16``, as the gradients of the
    16-bit GEMMs already are) halves the bytes on the wire; the mean is taken in float32 after the sum."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    world = torch.distributed.get_world_size(group)
    dt = comm_dtype or params[0].grad.dtype
    item = torch.empty((), dtype=dt).element_size()
    buckets, cur, cur_bytes = [], [], 0
    for p in params:
        nb = p.grad.numel() * item
        if cur and cur_bytes + nb > bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nb
    buckets.append(cur)
    pending = []
    for bucket in buckets:
        flat = torch.cat([p.grad.reshape(-1) for p in bucket]).to(dt)  # one batched gather + one cast per bucket
        pending.append((bucket, flat, torch.distributed.all_reduce(flat, group=group, async_op=True)))
    for bucket, flat, work in pending:
        work.wait()
        mean = flat.to(bucket[0].grad.dtype).div_(world)
        views, o = [], 0
        for p in bucket:
            n = p.grad.numel()
            views.append(mean[o:o + n].view_as(p.grad))
            o += n
        torch._foreach_copy_([p.grad for p in bucket], views)  # one multi-tensor scatter per bucket


def default_clip_hparams(distortion="lossy_Z", beta=5e-2, max_steps=1000, n_data=50000, clip_state_dict=None,
                         is_frozen=True, compute_dtype=torch.bfloat16):
    """hparams of the reference's CLIP bottleneck recipe (config/featurizer/bottleneck_clip_lossyZ.yaml,
    bin/clip/clip_hub.sh) with ``rate=H_factorized`` and either the lossy-Z or the BINCE
    (contrastive, config/distortion/BINCE.yaml) distortion.  ``is_frozen=False`` fine-tunes the CLIP tower
    as well (the reference's recipe without the ``Freezer`` callback of config/finetune/freezer.yaml)."""
    if distortion == "lossy_Z":
        dist = dict(mode="lossy_Z", factor_beta=1, kwargs=dict(p_norm=1))
    elif distortion == "BINCE":
        dist = dict(mode="contrastive", factor_beta=0.003,
                    kwargs=dict(is_train_temperature=True, temperature=0.01, effective_batch_size=n_data, is_project=True,
                                is_cosine=True, project_kwargs=dict(mode="mlp", out_shape=0.5, in_shape=512)))
    else:
        raise ValueError(distortion)
    return NamespaceMap(
        stage="feat",
        data=dict(shape=(3, 224, 224), mode="image", max_steps=max_steps, length=n_data),
        encoder=dict(z_dim=512, fam="deterministic", fam_kwargs={}, arch="clip", arch_kwargs=dict(clip_state_dict=clip_state_dict, is_frozen=is_frozen, compute_dtype=compute_dtype)),
        rate=dict(mode="H_factorized", factor_beta=1,
                  kwargs=dict(z_dim=512, is_endToEnd=False, warmup_k_epoch=0,
                              kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))),
        distortion=dist,
        featurizer=dict(is_features=True, loss=dict(beta=beta, beta_anneal="linear")),
        evaluation=dict(featurizer=dict(is_online=False)),
        optimizer_feat=dict(mode="AdamW", kwargs=dict(lr=1e-3, weight_decay=3e-8)),
        optimizer_coder=dict(mode="AdamW", kwargs=dict(lr=3e-4, weight_decay=1e-6)),
        scheduler_feat=None, scheduler_coder=None,
    )
