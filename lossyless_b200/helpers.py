"""Hot-path helpers of the reference's ``lossyless/helpers.py`` (the plotting half is out of
scope): ``BASE_LOG`` (:27), ``Timer`` (:45-56), ``atleast_ndim`` (:330), ``Delta`` (:338-392),
``Annealer`` (:768-870), ``GatherFromGpus`` (:873-893), ``weights_init`` (:153), optimiser /
scheduler factories (:492-593) and a minimal attribute-dict for hparams
(utils/helpers.py ``NamespaceMap``)."""
import time
from functools import reduce
from numbers import Number

import os

import torch
import torch.nn as nn
from torch.distributions import Distribution, constraints
from torch.distributions.utils import broadcast_all

BASE_LOG = 2  # rates are logged in bits, losses are computed in nats

MEANS = dict(clip=[0.48145466, 0.4578275, 0.40821073])
STDS = dict(clip=[0.26862954, 0.26130258, 0.27577711])


class Timer:
    """Context manager measuring wall time.  With ``sync=True`` CUDA is synchronised on both edges, so the duration is
    the device time of the enclosed work (the reference does not synchronise, which is only correct where a
    ``.tolist()`` inside the region does it for it).  ``sync=False`` is the reference's plain host timer: a training
    step must not drain the device twice for a duration nobody reads."""

    def __init__(self, sync=True):
        self.sync = sync and torch.cuda.is_available()

    def __enter__(self):
        if self.sync:
            torch.cuda.synchronize()
        self.start = time.perf_counter()
        return self

    def __exit__(self, *args):
        if self.sync:
            torch.cuda.synchronize()
        self.end = time.perf_counter()
        self.duration = self.end - self.start


def mean(array):
    return sum(array) / len(array)


def prod(iterable):
    return reduce(lambda a, b: a * b, iterable, 1)


def atleast_ndim(x, ndim):
    if x is None:
        return None
    return x.view(list(x.shape) + [1] * (ndim - x.ndim))


def batch_flatten(x):
    shape = x.shape
    return x.reshape(-1, shape[-1]), shape


def batch_unflatten(x, shape):
    return x.reshape(*shape[:-1], -1)


class OrderedSet(dict):
    """Insertion-ordered set of (hashable) parameters with the set algebra the reference uses."""

    def __init__(self, items=()):
        super().__init__((i, None) for i in items)

    def __or__(self, other):
        return OrderedSet(list(self) + [o for o in other if o not in self])

    def __sub__(self, other):
        other = set(other)
        return OrderedSet([i for i in self if i not in other])

    def __and__(self, other):
        other = set(other)
        return OrderedSet([i for i in self if i in other])


class NamespaceMap(dict):
    """dict with attribute access, nested (enough of utils/helpers.py:44-106 for hparams)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = NamespaceMap(v) if isinstance(v, dict) and not isinstance(v, NamespaceMap) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


def weights_init(module, nonlinearity="relu"):
    """Initialise direct children: Kaiming for conv / linear, recursion through reset_parameters."""
    for m in module.children():
        if isinstance(m, nn.modules.conv._ConvNd):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity=nonlinearity)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            nn.init.kaiming_uniform_(m.weight, nonlinearity=nonlinearity)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif hasattr(m, "reset_parameters"):
            m.reset_parameters()
        else:
            weights_init(m, nonlinearity=nonlinearity)


class Delta(Distribution):
    """Point mass at ``loc`` (rsample returns loc)."""

    has_rsample = True
    arg_constraints = {"loc": constraints.real, "log_density": constraints.real}
    support = constraints.real

    def __init__(self, loc, log_density=0.0, validate_args=False):
        # No argument validation and no host scalar on the way in: torch's default (`validate_args=None` -> check that
        # loc is real -> `.all()` read back by the host) and `broadcast_all(loc, 0.0)` (a pageable host-to-device copy
        # of the scalar) each synchronise the stream, once per encoder call and per `.detach()` of the training step.
        if torch.is_tensor(loc) and isinstance(log_density, Number):
            self.loc = loc
            self.log_density = loc.new_full((), float(log_density)).expand(loc.shape)
        else:
            self.loc, self.log_density = broadcast_all(loc, log_density)
        batch_shape = torch.Size() if isinstance(loc, Number) else self.loc.size()
        super().__init__(batch_shape, validate_args=validate_args)

    @property
    def mean(self):
        return self.loc

    @property
    def variance(self):
        return torch.zeros_like(self.loc)

    def rsample(self, sample_shape=torch.Size()):
        return self.loc.expand(list(sample_shape) + list(self.loc.shape))

    def log_prob(self, x):
        return (x == self.loc).type(x.dtype).log() + self.log_density


class Annealer:
    """Hyper-parameter annealing: linear / geometric / constant from ``initial_value`` to
    ``final_value`` over ``n_steps_anneal`` calls after ``start_step``."""

    def __init__(self, initial_value, final_value, n_steps_anneal, start_step=0, default=None, mode="geometric"):
        if n_steps_anneal < 0:
            n_steps_anneal *= -1
            initial_value, final_value = final_value, initial_value
        self.initial_value, self.final_value = initial_value, final_value
        self.n_steps_anneal, self.start_step = n_steps_anneal, start_step
        self.default = default if default is not None else self.initial_value
        self.mode = mode.lower()
        if self.mode == "linear":
            self.factor = (self.final_value - self.initial_value) / self.n_steps_anneal
        elif self.mode == "geometric":
            self.factor = (self.final_value / self.initial_value) ** (1 / self.n_steps_anneal)
        elif self.mode != "constant":
            raise ValueError(f"Unkown mode : {mode}.")
        self.reset_parameters()

    def reset_parameters(self):
        self.n_training_calls = 0

    def is_annealing(self, n_update_calls):
        return self.mode != "constant" and n_update_calls < (self.n_steps_anneal + self.start_step)

    def __call__(self, is_update=False, n_update_calls=None):
        if is_update:
            self.n_training_calls += 1
        if n_update_calls is None:
            n_update_calls = self.n_training_calls
        if self.start_step > n_update_calls:
            return self.default
        k = n_update_calls - self.start_step
        if self.is_annealing(n_update_calls):
            if self.mode == "geometric":
                return self.initial_value * self.factor ** k
            return self.initial_value + self.factor * k
        return self.final_value


class GatherFromGpus(torch.autograd.Function):
    """all_gather that keeps only the local gradient in backward (the reference's semantics:
    remote tiles are constants for this rank)."""

    @staticmethod
    def forward(ctx, tensor):
        ctx.rank = torch.distributed.get_rank()
        gathered = [torch.zeros_like(tensor) for _ in range(torch.distributed.get_world_size())]
        torch.distributed.all_gather(gathered, tensor.contiguous())
        return tuple(gathered)

    @staticmethod
    def backward(ctx, *grads):
        return grads[ctx.rank].clone()


gather_from_gpus = GatherFromGpus.apply


class FlatAdamW(torch.optim.AdamW):
    """``torch.optim.AdamW`` whose share of an unfrozen CLIP tower runs on the flat-buffer kernels (csrc/optim.cu).

    Constructed like AdamW with ALL the parameters of the optimiser (so ``param_groups``, ``lr`` and the schedulers
    behave as usual).  Parameters that belong to a ``clip_vit.VisionTransformer`` in fine-tuning mode are handled
    natively: the tower is switched to its flat training state (``FlatTrainState``), whose backward pass leaves the
    gradients in two flat buffers instead of ``p.grad``; ``step()`` then runs torch's fused AdamW on the remaining
    parameters (``p.grad is None`` skips the tower's) and ONE native AdamW launch per flat buffer, which also writes
    next step's 16-bit weights.  ``clip_grad_norm_`` is the global-norm clip over both kinds of gradients, computed on
    the device (no host synchronisation) and applied to the flat ones through the kernel's ``grad_scale`` operand --
    together with the 1 / world_size of the data-parallel mean, whose all-reduces (sum) were launched block by block
    from inside the backward pass."""

    def __init__(self, params, **kwargs):
        params = list(params)
        self._towers = []
        for p in params:
            ref = getattr(p, "_ll_vit", None)
            vit = ref() if ref is not None else None
            if vit is not None and vit._trainable and all(vit is not t for t in self._towers):
                self._towers.append(vit)
        super().__init__(params, **kwargs)
        for vit in self._towers:
            vit.enable_flat_training_(True)
        self._grad_scale = None

    def flat_states(self):
        return [f for f in (vit.flat_state() for vit in self._towers) if f is not None]

    def prepare_backward(self, sync=False, group=None, world=1):
        """Before ``loss.backward()``: whether (and over which process group) the towers all-reduce their gradient
        slices from inside the backward pass."""
        for f in self.flat_states():
            f.sync, f.sync_group, f.sync_world = bool(sync), group, int(world)
        self._world = int(world) if sync else 1
        self._grad_scale = None

    def finish_backward(self):
        for f in self.flat_states():
            f.wait_reductions()

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm):
        """Global-norm clipping over every gradient of this optimiser; returns the total norm (a device tensor)."""
        world = getattr(self, "_world", 1)
        others = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        dev = others[0].device if others else self.flat_states()[0].device
        total = torch.zeros(1, dtype=torch.float32, device=dev)
        if others:
            total += torch.stack(torch._foreach_norm(others)).float().square().sum()  # one square for all, not one launch per tensor
        for f in self.flat_states():
            if f.has_grads:
                sq = torch.empty(1, dtype=torch.float32, device=dev)
                total += f.sumsq(sq) / float(world * world)  # the flat gradients still hold the SUM over ranks
        norm = total.sqrt()
        coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        if others:
            torch._foreach_mul_(others, coef.reshape(()))
        self._grad_scale = (coef / float(world)).float().contiguous()
        return norm

    @torch.no_grad()
    def step(self, closure=None):
        loss = super().step(closure)
        g = self.param_groups[0]
        world = getattr(self, "_world", 1)
        scale = self._grad_scale
        if scale is None and world > 1:
            scale = torch.full((1,), 1.0 / world, dtype=torch.float32, device=self.flat_states()[0].device)
        for f in self.flat_states():
            if f.has_grads:
                f.adamw_step(g["lr"], g["betas"], g["eps"], g["weight_decay"], scale)
        self._grad_scale = None
        return loss


def get_optimizer(parameters, mode, **kwargs):
    Optimizer = getattr(torch.optim, mode)
    kwargs = dict(kwargs)
    if "lr_factor" in kwargs:
        kwargs["lr"] = kwargs["lr"] * kwargs.pop("lr_factor")
    parameters = list(parameters)
    if mode in ("Adam", "AdamW", "SGD") and "fused" not in kwargs and "foreach" not in kwargs and parameters \
            and all(p.is_cuda and torch.is_floating_point(p) for p in parameters):
        kwargs["fused"] = True  # same update rule, one multi-tensor kernel instead of ~10 per step
    if mode == "AdamW" and not kwargs.get("amsgrad", False) and os.environ.get("LL_FLAT_OPTIMIZER", "1") != "0" \
            and any(getattr(p, "_ll_vit", None) is not None for p in parameters):
        return FlatAdamW(parameters, **kwargs)  # an unfrozen CLIP tower among the parameters: flat-buffer kernels
    return Optimizer(parameters, **kwargs)


def get_lr_scheduler(optimizer, scheduler_type, epochs=None, decay_factor=1000, k_steps=3, name=None, **kwargs):
    if scheduler_type is None:
        return None
    if scheduler_type == "expdecay":
        return torch.optim.lr_scheduler.ExponentialLR(optimizer, (1 / decay_factor) ** (1 / epochs))
    if scheduler_type == "UniformMultiStepLR":
        delta = epochs // (k_steps + 1)
        return torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=[delta * i for i in range(1, k_steps + 1)],
                                                    gamma=(1 / decay_factor) ** (1 / k_steps))
    return getattr(torch.optim.lr_scheduler, scheduler_type)(optimizer, **kwargs)


def append_optimizer_scheduler_(hparams_opt, hparams_sch, parameters, optimizers, schedulers, name=None):
    params = [p for p in parameters if p.requires_grad]
    optimizer = get_optimizer(params, hparams_opt.mode, **hparams_opt.kwargs)
    optimizers.append(optimizer)
    for mode in (hparams_sch.modes if hparams_sch is not None else []):
        sch = get_lr_scheduler(optimizer, mode, name=name, **hparams_sch.kwargs.get(mode, {}))
        if sch is not None:
            schedulers.append(dict(scheduler=sch, name=name))
    return optimizers, schedulers
