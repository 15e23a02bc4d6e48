"""Rate estimators on the B200 kernels (reference lossyless/rates.py).

On the hot path (SURVEY.md section 8a-4/4'/5/6): ``get_rate_estimator`` (:29-49), the 2-D
``EntropyBottleneck`` wrapper (:55-71), ``RateEstimator`` (:75-314), ``HRateEstimator``
(:398-506) and ``HRateFactorizedPrior`` (:509-564).  ``forward`` is one fused launch
(process_z_in -> noise/round -> likelihood -> -log q row sum -> process_z_out) with a fused
backward; ``compress`` / ``decompress`` are the batched rANS kernels.  ``HRateHyperprior`` (:571-756, the
first "next" row of section 8f) is built on the same kernels (side bottleneck + indexed rANS for the
Gaussian conditional).  ``lossless``, ``H_spatial`` and ``MI`` raise NotImplementedError.
"""
import math

import torch
import torch.nn as nn

from . import entropy_models as em
from .helpers import BASE_LOG, OrderedSet, Timer, atleast_ndim, mean, weights_init

__all__ = ["get_rate_estimator", "EntropyBottleneck", "RateEstimator", "HRateEstimator", "HRateFactorizedPrior",
           "HRateHyperprior", "get_scale_table"]


def get_rate_estimator(mode, z_dim=None, p_ZlX=None, **kwargs):
    """Return the correct entropy coder."""
    if mode == "H_factorized":
        return HRateFactorizedPrior(z_dim, **kwargs)
    if mode == "H_hyper":
        return HRateHyperprior(z_dim, **kwargs)
    if mode in ("lossless", "H_spatial", "MI"):
        raise NotImplementedError(f"rate estimator {mode!r} is a 'next' row of SURVEY.md section 8f; H_factorized and H_hyper are built")
    raise ValueError(f"Unkown rate estimator={mode}.")


class EntropyBottleneck(em.EntropyBottleneck):
    """The reference's wrapper: treats z as a vector [B, C] instead of a [B, C, 1, 1] image."""

    def forward(self, z, noise=None):
        return super().forward(z, noise=noise)

    def compress(self, z, **kw):
        return super().compress(atleast_ndim(z, 4), **kw)

    def decompress(self, z_strings, **kw):
        z_hat = super().decompress(z_strings, [1, 1], **kw)
        return z_hat.reshape(z_hat.shape[0], -1)


class RateEstimator(nn.Module):
    """Base class of coders: ``forward(z, p_Zlx, parent) -> (z_hat, rates [nats], logs, other)``."""

    is_can_compress = False

    def __init__(self, z_dim, warmup_k_epoch=0, is_endToEnd=True):
        super().__init__()
        self.z_dim, self.warmup_k_epoch, self.is_endToEnd = z_dim, warmup_k_epoch, is_endToEnd

    def reset_parameters(self):
        weights_init(self)

    def forward(self, z, p_Zlx, parent=None):
        with torch.autocast("cuda", enabled=False):  # precision here is important
            z_hat, rates, r_logs, r_other = self.forward_help(z, p_Zlx, parent)
            epoch = getattr(parent, "current_epoch", 0) if parent is not None else 0
            if (not self.is_endToEnd) or (epoch < self.warmup_k_epoch):
                # disjoint training: z_hat is still returned, but the rate only trains the coder
                z_detached = z.detach() + z * 0
                p_detached = p_Zlx.detach(is_grad_flow=True) if p_Zlx is not None else None
                _, rates, *_ = self.forward_help(z_detached, p_detached, parent)
        return z_hat, rates, r_logs, r_other

    def forward_help(self, z, p_Zlx, parent=None):
        raise NotImplementedError()

    def compress(self, z, parent=None):
        raise NotImplementedError()

    def decompress(self, all_strings):
        raise NotImplementedError()

    def real_rate(self, z, is_return_logs=False, parent=None):
        """Actual bits per image after entropy coding (mean over the batch)."""
        batch, z_dim = z.shape
        with Timer() as compress_timer:
            all_strings = self.compress(z, parent=parent)
        if is_return_logs:
            with Timer() as decompress_timer:
                _ = self.decompress(all_strings)
        n_bits = 8 * sum(mean([len(s) for s in strings]) for strings in all_strings)
        if is_return_logs:
            return n_bits, dict(compress_time=compress_timer.duration / batch,
                                receiver_time=decompress_timer.duration / batch, n_bits=n_bits)
        return n_bits

    def aux_loss(self):
        raise NotImplementedError()

    def aux_parameters(self):
        raise NotImplementedError()

    def make_pickable_(self):
        for m in self.modules():
            if isinstance(m, em.EntropyModel):
                m.entropy_coder = None

    def undo_pickable_(self):
        for m in self.modules():
            if isinstance(m, em.EntropyModel):
                m.entropy_coder = em._RansCoder()

    def update(self, force=False):
        updated = True
        for m in self.children():
            if isinstance(m, em.EntropyBottleneck):
                updated &= m.update(force=force)
            elif isinstance(m, em.GaussianConditional):
                updated &= m.update_scale_table(get_scale_table().tolist(), force=force)
        return updated

    def prepare_compressor_(self):
        """Make sure the model can be used for actual compression (coder handle + CDF tables)."""
        self.undo_pickable_()
        self.update(force=True)


class HRateEstimator(RateEstimator):
    """Entropy coders with a learned per-dimension affine map in front (scaling / biasing)."""

    is_can_compress = True

    def __init__(self, z_dim, kwargs_ent_bottleneck={}, **kwargs):
        super().__init__(z_dim, **kwargs)
        self.kwargs_ent_bottleneck = kwargs_ent_bottleneck
        self.scaling = torch.nn.Parameter(torch.ones(z_dim))
        self.biasing = torch.nn.Parameter(torch.zeros(z_dim))
        self._es_cache = None

    def aux_loss(self):
        """compressai CompressionModel.aux_loss: sum of the bottlenecks' quantile losses."""
        return sum(m.loss() for m in self.modules() if isinstance(m, em.EntropyBottleneck))

    def reset_parameters(self):
        weights_init(self)
        self.scaling = torch.nn.Parameter(torch.ones(self.z_dim))
        self.biasing = torch.nn.Parameter(torch.zeros(self.z_dim))

    def process_z_in(self, z):
        return (z.float() + self.biasing) * self.scaling.exp()

    def process_z_out(self, z_hat):
        return (z_hat / self.scaling.exp()) - self.biasing

    def codec_exp_scale(self):
        """exp(scaling) for the integer path, evaluated in float32 on the CPU (device independent)."""
        key = (self.scaling._version, self.scaling.data_ptr(), str(self.scaling.device))
        if self._es_cache is None or self._es_cache[0] != key:
            self._es_cache = (key, self.scaling.detach().float().cpu().exp().to(self.scaling.device))
        return self._es_cache[1]

    def aux_parameters(self):
        return OrderedSet(p for n, p in self.named_parameters() if n.endswith(".quantiles"))

    @property
    def is_coder_updated(self):
        return all(m._offset.numel() > 0 for m in self.modules() if isinstance(m, em.EntropyModel))

    @property
    def is_coder_present(self):
        return all(m.entropy_coder is not None for m in self.modules() if isinstance(m, em.EntropyModel))

    @property
    def is_compute_real_rate(self):
        return (not self.training) and self.is_coder_updated and self.is_coder_present


class HRateFactorizedPrior(HRateEstimator):
    """Codes with the (approximate) entropy H[Z] of a factorised prior (Balle et al. 2018)."""

    def __init__(self, z_dim, **kwargs):
        super().__init__(z_dim, **kwargs)
        self.entropy_bottleneck = EntropyBottleneck(self.z_dim, **self.kwargs_ent_bottleneck)
        self.reset_parameters()

    def reset_parameters(self):
        # weights_init would recurse into the bottleneck; it keeps its own (compressai) init
        self.scaling = torch.nn.Parameter(torch.ones(self.z_dim))
        self.biasing = torch.nn.Parameter(torch.zeros(self.z_dim))

    def forward_help(self, z, _, parent=None, noise=None):
        nz = None
        if self.training:
            nz = noise if noise is not None else torch.empty_like(z, dtype=torch.float32).uniform_(-0.5, 0.5)
        # fused: process_z_in -> bottleneck -> (-log q).sum(-1) -> process_z_out
        z_hat, q_z, neg_log_q_z = em.eb_rate_forward(self.entropy_bottleneck, z, nz, self.biasing, self.scaling)
        logs = dict(H_q_Z=neg_log_q_z.mean() / math.log(BASE_LOG), H_ZlX=0)
        if self.is_compute_real_rate:
            n_bits, logs2 = self.real_rate(z, is_return_logs=True)
            logs.update(logs2)
            logs["n_bits"] = n_bits
        return z_hat, neg_log_q_z, logs, dict()

    def compress(self, z, parent=None):
        # list for generality when hyperprior; process_z_in is fused into the encode kernel
        return [self.entropy_bottleneck.compress(z.detach(), bias=self.biasing, exp_scale=self.codec_exp_scale())]

    def decompress(self, all_strings):
        assert isinstance(all_strings, list) and len(all_strings) == 1
        return self.entropy_bottleneck.decompress(all_strings[0], bias=self.biasing, exp_scale=self.codec_exp_scale())


def get_scale_table(min=0.11, max=256, levels=64):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class HRateHyperprior(HRateEstimator):
    """Scale (and mean) hyperprior rate (Balle et al. 2018; Minnen et al. 2018), reference
    lossyless/rates.py:571-756: side information ``s = side_encoder(z_in)`` is coded with a factorised
    bottleneck, ``z_encoder(s_hat)`` predicts per-dimension Gaussian scales (and means) under which z is
    coded; the rate is ``-log q(s) - log q(z|s)``.  Two rANS streams per image.

    The reference's compress path overwrites ``means_hat`` with ``scales_hat`` (rates.py:691-692), which
    makes its real bit-rate far larger than its entropy estimate; that defect is not reproduced: the
    predicted means are used for coding exactly as in ``forward_help``."""

    def __init__(self, z_dim, factor_dim=5, side_z_dim=None, is_pred_mean=True, **kwargs):
        super().__init__(z_dim, **kwargs)
        if side_z_dim is None:
            side_z_dim = max(10, self.z_dim // factor_dim)
        self.side_z_dim, self.is_pred_mean = side_z_dim, is_pred_mean
        self.entropy_bottleneck = EntropyBottleneck(side_z_dim, **self.kwargs_ent_bottleneck)
        self.gaussian_conditional = em.GaussianConditional(None)
        self.side_encoder, self.z_encoder = self.get_encoders()
        self.reset_parameters()

    def reset_parameters(self):
        self.side_encoder.reset_parameters()
        self.z_encoder.reset_parameters()
        self.scaling = torch.nn.Parameter(torch.ones(self.z_dim))
        self.biasing = torch.nn.Parameter(torch.zeros(self.z_dim))

    def get_encoders(self):
        from .architectures import MLP

        kwargs_mlp = dict(n_hid_layers=2, hid_dim=max(self.z_dim, 256))
        side_encoder = MLP(self.z_dim, self.side_z_dim, **kwargs_mlp)
        z_encoder = MLP(self.side_z_dim, self.z_dim * (2 if self.is_pred_mean else 1), **kwargs_mlp)
        return side_encoder, z_encoder

    def chunk_params(self, gaussian_params):
        if self.is_pred_mean:
            return gaussian_params.chunk(2, -1)
        return gaussian_params, None

    def forward_help(self, z, _, __, noise=None):
        z_in = self.process_z_in(z)
        side_z = self.side_encoder(z_in)
        nz_side = noise[1] if noise is not None else None
        side_z_hat, q_s = self.entropy_bottleneck(side_z, noise=nz_side)
        scales_hat, means_hat = self.chunk_params(self.z_encoder(side_z_hat))
        # Gaussian-conditional likelihood and its per-row -log q in one fused launch (csrc/gaussian.cu)
        z_hat, q_zls, neg_log_q_zls = self.gaussian_conditional.forward_rates(z_in, scales_hat, means=means_hat,
                                                                                noise=noise[0] if noise is not None else None)
        neg_log_q_s = -torch.log(q_s).sum(-1)
        neg_log_q_zs = neg_log_q_s + neg_log_q_zls
        logs = dict(H_q_ZlS=neg_log_q_zls.mean() / math.log(BASE_LOG), H_q_Z=neg_log_q_zs.mean() / math.log(BASE_LOG),
                    H_q_S=neg_log_q_s.mean() / math.log(BASE_LOG), H_ZlX=0)
        if self.is_compute_real_rate:
            n_bits, logs2 = self.real_rate(z, is_return_logs=True)
            logs.update(logs2)
            logs["n_bits"] = n_bits
        return self.process_z_out(z_hat), neg_log_q_zs, logs, dict()

    def get_indexes_means_hat(self, side_z_strings):
        side_z_hat = self.entropy_bottleneck.decompress(side_z_strings)
        scales_hat, means_hat = self.chunk_params(self.z_encoder(side_z_hat))
        return self.gaussian_conditional.build_indexes(scales_hat), means_hat

    @torch.no_grad()
    def compress(self, z, parent=None):
        z_in = self.process_z_in(z)
        side_z_strings = self.entropy_bottleneck.compress(self.side_encoder(z_in))
        indexes, means_hat = self.get_indexes_means_hat(side_z_strings)
        z_strings = self.gaussian_conditional.compress(z_in, indexes, means=means_hat)
        return [z_strings, side_z_strings]

    @torch.no_grad()
    def decompress(self, all_strings):
        assert isinstance(all_strings, list) and len(all_strings) == 2
        z_strings, side_z_strings = all_strings
        indexes, means_hat = self.get_indexes_means_hat(side_z_strings)
        z_hat = self.gaussian_conditional.decompress(z_strings, indexes, means=means_hat)
        return self.process_z_out(z_hat)
