"""Conditional-distribution wrapper of the encoder (reference lossyless/distributions.py:18-153):
``CondDist`` maps x to a torch Distribution over z; the hot path uses the ``deterministic``
family (point mass, ``rsample()`` returns the encoder output)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Independent, Normal

from .helpers import Delta, weights_init

__all__ = ["CondDist", "Deterministic", "DiagGaussian"]


class CondDist(nn.Module):
    """``CondDist(in_shape, out_dim, Architecture, family)``: ``mapper = Architecture(in_shape,
    out_dim * n_param)``; ``forward(x)`` returns ``Family.from_suff_param(mapper(x))``."""

    def __init__(self, in_shape, out_dim, Architecture, family, **kwargs):
        super().__init__()
        if family == "diaggaussian":
            self.Family = DiagGaussian
        elif family == "deterministic":
            self.Family = Deterministic
        else:
            raise ValueError(f"Unkown family={family}.")
        self.in_shape, self.out_dim, self.kwargs = in_shape, out_dim, kwargs
        self.mapper = Architecture(in_shape, out_dim * self.Family.n_param)
        self.reset_parameters()

    def forward(self, x):
        return self.Family.from_suff_param(self.mapper(x), **self.kwargs)

    def forward_many(self, xs):
        """``[forward(x) for x in xs]`` through ONE mapper call on the concatenated batch.  Only valid for
        mappers that treat samples independently (no batch statistics): the caller checks."""
        xs = list(xs)
        if hasattr(self.mapper, "forward_views"):
            out = self.mapper.forward_views(xs)  # no concatenated copy of the (float32) views
        else:
            out = self.mapper(torch.cat(xs, dim=0))
        sizes = [x.shape[0] for x in xs]
        return [self.Family.from_suff_param(o, **self.kwargs) for o in out.split(sizes, dim=0)]

    def reset_parameters(self):
        weights_init(self)


class _Family:
    n_param = None

    @classmethod
    def from_suff_param(cls, concat, **kwargs):
        params = concat.reshape(concat.shape[0], -1, cls.n_param).unbind(-1)
        return cls(*cls.preprocess_suff_params(*params), **kwargs)

    @classmethod
    def preprocess_suff_params(cls, *params):
        return params


class Deterministic(_Family, Independent):
    """Delta distribution: no stochasticity."""

    n_param = 1

    def __init__(self, param):
        super().__init__(Delta(param), 1)

    def detach(self, is_grad_flow=False):
        loc = self.base_dist.loc.detach()
        if is_grad_flow:
            loc = loc + 0 * self.base_dist.loc
        return Deterministic(loc)


class DiagGaussian(_Family, Independent):
    n_param = 2
    min_std = 1e-5

    def __init__(self, diag_loc, diag_scale):
        super().__init__(Normal(diag_loc, diag_scale), 1)

    @classmethod
    def preprocess_suff_params(cls, diag_loc, diag_log_var):
        return diag_loc, F.softplus(diag_log_var) + cls.min_std

    def detach(self, is_grad_flow=False):
        loc, scale = self.base_dist.loc.detach(), self.base_dist.scale.detach()
        if is_grad_flow:
            loc, scale = loc + 0 * self.base_dist.loc, scale + 0 * self.base_dist.scale
        return DiagGaussian(loc, scale)
