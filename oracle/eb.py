"""Oracle restatement of compressai==1.1.5 ``EntropyBottleneck`` (factorised prior).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).  PARITY UNPINNED.

Follows compressai/entropy_models/entropy_models.py (``EntropyBottleneck._logits_cumulative``,
``_likelihood``, ``forward``, ``update``, ``loss``; ``EntropyModel.quantize/dequantize/
_pmf_to_cdf``) as called from the reference at hub/compressor.py:49-63,98-100,124 and
lossyless/rates.py:55-71,426,534-564.

Two arithmetic back-ends:
  * ``f32``  -- torch CPU float32 ops in exactly the op order of the original (the original
    *is* torch code), used to derive the integer CDF tables and golden vectors;
  * ``f64``  -- numpy float64, used only as an accuracy yard-stick for tolerances.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import rans

FILTERS = (3, 3, 3, 3)
LIKELIHOOD_BOUND = 1e-9
TAIL_MASS = 1e-9


def split_state_dict(sd, prefix="entropy_bottleneck."):
    """Pull the EntropyBottleneck tensors (float32, CPU) out of a hub state dict."""
    n = len(FILTERS) + 1
    eb = {
        "matrices": [sd[f"{prefix}_matrix{i}"].detach().float().cpu() for i in range(n)],
        "biases": [sd[f"{prefix}_bias{i}"].detach().float().cpu() for i in range(n)],
        "factors": [sd[f"{prefix}_factor{i}"].detach().float().cpu() for i in range(n - 1)],
        "quantiles": sd[f"{prefix}quantiles"].detach().float().cpu(),
    }
    return eb


# --------------------------------------------------------------------- float32 (torch)
def logits_cumulative(eb, x):
    """x: [C,1,N] -> [C,1,N].  compressai EntropyBottleneck._logits_cumulative."""
    logits = x
    n = len(eb["matrices"])
    for i in range(n):
        logits = torch.matmul(F.softplus(eb["matrices"][i]), logits)
        logits = logits + eb["biases"][i]
        if i < n - 1:
            logits = logits + torch.tanh(eb["factors"][i]) * torch.tanh(logits)
    return logits


def likelihood(eb, v):
    """v: [C,1,N] (already noised / dequantised) -> likelihood before the lower bound."""
    lower = logits_cumulative(eb, v - 0.5)
    upper = logits_cumulative(eb, v + 0.5)
    sign = -torch.sign(lower + upper)
    return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))


def medians(eb):
    return eb["quantiles"][:, 0, 1].clone()


class _LowerBound(torch.autograd.Function):
    """compressai/ops/bound_ops.py LowerBound: max(x, bound); the gradient passes where
    x >= bound or where it would increase x (grad < 0)."""

    @staticmethod
    def forward(ctx, x, bound):
        b = torch.full_like(x, bound)
        ctx.save_for_backward(x, b)
        return torch.max(x, b)

    @staticmethod
    def backward(ctx, g):
        x, b = ctx.saved_tensors
        return ((x >= b) | (g < 0)).type_as(g) * g, None


def forward(eb, z_in, noise=None):
    """compressai EntropyBottleneck.forward on a [B,C] input (the reference's 2-D wrapper,
    lossyless/rates.py:55-62).  ``noise`` [B,C] in (-.5,.5) selects training mode (the
    U(-1/2,1/2) sample is injected so the test is deterministic); None = eval mode
    (round(x - median) + median).  Returns (z_hat [B,C], q_z [B,C]) with LowerBound(1e-9)."""
    x = z_in.float().t().contiguous().unsqueeze(1)  # [C,1,B]
    if noise is not None:
        v = x + noise.float().t().contiguous().unsqueeze(1)
    else:
        med = eb["quantiles"][:, :, 1:2]
        v = torch.round(x - med) + med
    lik = likelihood(eb, v)
    lik = _LowerBound.apply(lik, LIKELIHOOD_BOUND)
    return v[:, 0, :].t().contiguous(), lik[:, 0, :].t().contiguous()


def aux_loss(eb, target=None):
    """compressai EntropyBottleneck.loss(): sum |logits_cumulative(quantiles) - target|."""
    if target is None:
        t = math.log(2 / TAIL_MASS - 1)
        target = torch.tensor([-t, 0.0, t])
    return torch.abs(logits_cumulative(eb, eb["quantiles"]) - target).sum()


def update(eb):
    """compressai EntropyBottleneck.update(): integer CDF tables.

    Returns dict(cdf int32[C, Lmax+2], cdf_length int32[C], offset int32[C], pmf f32[C,Lmax],
    tail f32[C], pmf_length int32[C])."""
    q = eb["quantiles"]
    med = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
    offset = -minima
    pmf_start = med - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max().item())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
    lower = logits_cumulative(eb, samples - 0.5)
    upper = logits_cumulative(eb, samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])  # global last col
    C = pmf.shape[0]
    cdf = np.zeros((C, max_length + 2), dtype=np.int32)
    pmf_np, tail_np = pmf.numpy(), tail.numpy()
    for c in range(C):
        L = int(pmf_length[c])
        prob = np.concatenate([pmf_np[c, :L], tail_np[c]])
        cdf[c, : L + 2] = rans.pmf_to_quantized_cdf(prob, 16)
    return dict(cdf=cdf, cdf_length=(pmf_length + 2).numpy().astype(np.int32),
                offset=offset.numpy().astype(np.int32), pmf=pmf_np, tail=tail_np[:, 0],
                pmf_length=pmf_length.numpy().astype(np.int32))


def process_z_in(z, biasing, scaling):
    """hub/compressor.py:105-109 / lossyless/rates.py:434-435 (without the 4-D unsqueeze)."""
    return (z.float() + biasing) * scaling.exp()


def process_z_out(z_hat, biasing, scaling):
    """hub/compressor.py:111-115 / lossyless/rates.py:437-438."""
    return (z_hat / scaling.exp()) - biasing


def compress(eb, tables, z_in):
    """compressai EntropyBottleneck.compress on [B,C] z_in -> list[bytes].
    symbols = int(round(z_in - median)) (EntropyModel.quantize 'symbols')."""
    med = eb["quantiles"][:, 0, 1]
    sym = torch.round(z_in.float() - med[None, :]).int().numpy()
    return rans.encode_batch(sym, tables["cdf"], tables["cdf_length"], tables["offset"]), sym


def decompress(eb, tables, strings):
    """compressai EntropyBottleneck.decompress -> z_hat_in [B,C] float32 (= symbols + median)."""
    C = tables["cdf"].shape[0]
    sym = rans.decode_batch(strings, C, tables["cdf"], tables["cdf_length"], tables["offset"])
    med = eb["quantiles"][:, 0, 1]
    return torch.from_numpy(sym).float() + med[None, :], sym


# --------------------------------------------------------------------- float64 (numpy)
def _softplus64(x):
    return np.where(x > 30, x, np.log1p(np.exp(np.minimum(x, 30))))


def logits_cumulative_f64(eb, x):
    """x: [C,N] float64 -> [C,N].  Same maths in float64, explicit small matmuls."""
    logits = x[:, None, :].astype(np.float64)  # [C,1,N]
    n = len(eb["matrices"])
    for i in range(n):
        M = _softplus64(eb["matrices"][i].numpy().astype(np.float64))  # [C,fo,fi]
        logits = np.einsum("coi,cin->con", M, logits) + eb["biases"][i].numpy().astype(np.float64)
        if i < n - 1:
            logits = logits + np.tanh(eb["factors"][i].numpy().astype(np.float64)) * np.tanh(logits)
    return logits[:, 0, :]


def likelihood_f64(eb, v):
    lower = logits_cumulative_f64(eb, v - 0.5)
    upper = logits_cumulative_f64(eb, v + 0.5)
    sign = -np.sign(lower + upper)
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))
    return np.abs(sig(sign * upper) - sig(sign * lower))
