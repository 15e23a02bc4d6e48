"""Oracle restatement of compressai==1.1.5 ``GaussianConditional`` and of the reference's scale-hyperprior
rate (lossyless/rates.py:567-756, with the defect at :691-692 -- ``means_hat`` overwritten by
``scales_hat`` -- NOT reproduced: the means predicted by ``z_encoder`` are used, as in ``forward_help``).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).  PARITY UNPINNED (compressai not
installable); float32 torch CPU ops in compressai's op order, integer work through oracle/rans_oracle.c.
"""
import math

import numpy as np
import scipy.stats
import torch

from . import rans

SCALE_BOUND = 0.11
TAIL_MASS = 1e-9
LIKELIHOOD_BOUND = 1e-9


def get_scale_table(min=0.11, max=256, levels=64):
    """lossyless/rates.py:567-568."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def standardized_cumulative(x):
    return 0.5 * torch.erfc(-(2 ** -0.5) * x)


def update(scale_table):
    """GaussianConditional.update(): integer CDF tables for every scale level."""
    scale_table = scale_table.float()
    multiplier = -scipy.stats.norm.ppf(TAIL_MASS / 2)
    pmf_center = torch.ceil(scale_table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(torch.max(pmf_length).item())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    samples_scale = scale_table.unsqueeze(1).float()
    upper = standardized_cumulative((0.5 - samples) / samples_scale)
    lower = standardized_cumulative((-0.5 - samples) / samples_scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    n = len(pmf_length)
    cdf = np.zeros((n, max_length + 2), dtype=np.int32)
    pmf_np, tail_np = pmf.numpy(), tail.numpy()
    for i in range(n):
        L = int(pmf_length[i])
        cdf[i, : L + 2] = rans.pmf_to_quantized_cdf(np.concatenate([pmf_np[i, :L], tail_np[i]]), 16)
    return dict(cdf=cdf, cdf_length=(pmf_length + 2).numpy().astype(np.int32), offset=(-pmf_center).numpy().astype(np.int32))


def lower_bound(x, bound):
    return torch.max(x, torch.tensor(bound, dtype=x.dtype))


def build_indexes(scales, scale_table):
    scales = lower_bound(scales, SCALE_BOUND)
    indexes = torch.full(scales.shape, len(scale_table) - 1, dtype=torch.int32)
    for s in scale_table[:-1]:
        indexes -= (scales <= s).int()
    return indexes


def forward(x, scales, means, noise=None):
    """GaussianConditional.forward -> (outputs, likelihood with LowerBound(1e-9))."""
    if noise is not None:
        out = x + noise
    else:
        out = torch.round(x - means) + means
    values = torch.abs(out - means)
    s = lower_bound(scales, SCALE_BOUND)
    lik = standardized_cumulative((0.5 - values) / s) - standardized_cumulative((-0.5 - values) / s)
    return out, lower_bound(lik, LIKELIHOOD_BOUND)


def compress(x, indexes, means, tables):
    """EntropyModel.compress with per-element indexes / means -> (list[bytes], symbols)."""
    sym = torch.round(x.float() - means).int().numpy()
    idx = indexes.numpy().astype(np.int32)
    out = [rans.encode_with_indexes(sym[r], idx[r], tables["cdf"], tables["cdf_length"], tables["offset"]) for r in range(sym.shape[0])]
    return out, sym


def decompress(strings, indexes, means, tables):
    idx = indexes.numpy().astype(np.int32)
    sym = np.stack([rans.decode_with_indexes(s, idx[r], tables["cdf"], tables["cdf_length"], tables["offset"]) for r, s in enumerate(strings)])
    return torch.from_numpy(sym).float() + means, sym
