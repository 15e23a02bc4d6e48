"""Distortion estimators on the B200 kernels (reference lossyless/distortions.py).

On the hot path: ``ContrastiveDistortion`` (:175-405) -- projector,
L2-normalise + logits tile kernel, all_gather of logit tiles with the reference's multi-rank
quirk, fused masked InfoNCE kernel (no materialised masked copy) -- and ``LossyZDistortion``
(:408-450).  ``DirectDistortion`` (:41-172) is kept for the CPU plumbing config (banana) in its
vector ("distribution") mode with plain torch ops; image decoders are out of scope.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from .architectures import get_Architecture
from .helpers import BASE_LOG, gather_from_gpus, weights_init

__all__ = ["get_distortion_estimator", "ContrastiveDistortion", "LossyZDistortion", "DirectDistortion", "prediction_loss"]


def get_distortion_estimator(mode, **kwargs):
    if mode == "direct":
        return DirectDistortion(**kwargs)
    elif mode == "contrastive":
        return ContrastiveDistortion(**kwargs)
    elif mode == "lossy_Z":
        return LossyZDistortion(**kwargs)
    raise ValueError(f"Unkown disotrtion.mode={mode}.")


# ------------------------------------------------------------------------------ kernels
class _LogitsTile(torch.autograd.Function):
    """zs [n, d] -> normalize(zs) @ normalize(zs).T, written into a [n, ld] buffer so remote tiles
    can be gathered to its right without a concat copy."""

    @staticmethod
    def forward(ctx, zs, n_cols):
        N.require_cuda(zs, "zs")
        zs = zs.contiguous().float()
        n, d = zs.shape
        zn = torch.empty_like(zs)
        inv = torch.empty(n, dtype=torch.float32, device=zs.device)
        logits = torch.empty((n, n_cols), dtype=torch.float32, device=zs.device)
        N.call("ll_contrastive_logits", N.ptr(zs), n, d, N.ptr(zn), N.ptr(inv), N.ptr(logits), n_cols, N.stream())
        ctx.save_for_backward(zn, inv)
        return logits

    @staticmethod
    def backward(ctx, g):
        zn, inv = ctx.saved_tensors
        n, d = zn.shape
        g = g.contiguous()
        g_zs = torch.empty_like(zn)
        N.call("ll_contrastive_logits_backward", N.ptr(g), g.shape[1], N.ptr(zn), N.ptr(inv), n, d, N.ptr(g_zs), N.stream())
        return g_zs, None


class _InfoNCE(torch.autograd.Function):
    """cross_entropy(masked(logits) / T, positives) over [n, n_cols] (own tile first)."""

    @staticmethod
    def forward(ctx, logits, inv_temp, pos_add):
        n, n_cols = logits.shape
        loss = torch.empty(n, dtype=torch.float32, device=logits.device)
        lse = torch.empty(n, dtype=torch.float32, device=logits.device)
        # 1 / temperature stays on the device (a float() here would make the host wait for the whole encoder
        # forward in the middle of every training step)
        it = inv_temp.detach().float().reshape(1).contiguous()
        N.call("ll_contrastive_ce_forward_dev", N.ptr(logits), n, n_cols, logits.stride(0), pos_add, N.ptr(it), N.ptr(loss), N.ptr(lse), N.stream())
        ctx.save_for_backward(logits, lse, it)
        ctx.pos_add = pos_add
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        logits, lse, it = ctx.saved_tensors
        n, n_cols = logits.shape
        g = torch.zeros_like(logits)  # remote tiles get no gradient (lossyless/helpers.py:885-890)
        g_it = torch.zeros(1, dtype=torch.float32, device=logits.device)
        N.call("ll_contrastive_ce_backward_dev", N.ptr(logits), n, n_cols, logits.stride(0), ctx.pos_add, N.ptr(it), N.ptr(lse),
               N.ptr(g_loss.contiguous().float()), N.ptr(g), g.stride(0), N.ptr(g_it), N.stream())
        return g, g_it.reshape(()), None


class _LossyZ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_hat, z, eps):
        N.require_cuda(z_hat, "z_hat")
        z_hat, z = z_hat.contiguous().float(), z.contiguous().float()
        rows, C = z_hat.shape
        dist = torch.empty(rows, dtype=torch.float32, device=z_hat.device)
        N.call("ll_lossy_z_forward", N.ptr(z_hat), N.ptr(z), rows, C, eps, N.ptr(dist), N.stream())
        ctx.save_for_backward(z_hat, z)
        ctx.eps = eps
        return dist

    @staticmethod
    def backward(ctx, g):
        z_hat, z = ctx.saved_tensors
        rows, C = z_hat.shape
        gz = torch.empty_like(z_hat)
        N.call("ll_lossy_z_backward", N.ptr(z_hat), N.ptr(z), rows, C, ctx.eps, N.ptr(g.contiguous().float()), N.ptr(gz), N.stream())
        return gz, -gz, None


# ------------------------------------------------------------------------------ modules
class ContrastiveDistortion(nn.Module):
    """Contrastive (InfoNCE) distortion with positives ``x_pos`` on the same orbit.

    Parameters follow the reference: ``temperature`` (lower bound when trained), ``is_train_temperature``
    (CLIP-style ``logit_scale``), ``is_cosine``, ``effective_batch_size`` (re-weights the negatives as if
    the batch were the whole dataset), ``is_already_featurized``, ``is_project`` + ``project_kwargs``."""

    def __init__(self, temperature=0.01, is_train_temperature=True, is_cosine=True, effective_batch_size=None,
                 is_already_featurized=False, is_project=True,
                 project_kwargs={"mode": "mlp", "out_shape": 128, "in_shape": 128}):
        super().__init__()
        self.temperature, self.is_train_temperature, self.is_cosine = temperature, is_train_temperature, is_cosine
        self.effective_batch_size, self.is_already_featurized, self.is_project = effective_batch_size, is_already_featurized, is_project
        if not is_cosine:
            raise NotImplementedError("the logits kernel implements the recommended is_cosine=True only")
        if self.is_project:
            project_kwargs = dict(project_kwargs)
            if project_kwargs["out_shape"] <= 1:
                project_kwargs["out_shape"] = max(10, int(project_kwargs["in_shape"] * project_kwargs["out_shape"]))
            self.projector = get_Architecture(**project_kwargs)()
        else:
            self.projector = torch.nn.Identity()
        if self.is_train_temperature:
            self.logit_scale = nn.Parameter(torch.log(torch.tensor(1 / 0.07)))  # same initialisation as CLIP
        self.reset_parameters()

    def reset_parameters(self):
        weights_init(self)
        if self.is_train_temperature:
            self.logit_scale = nn.Parameter(torch.log(torch.tensor(1 / 0.07)))

    def forward(self, z_hat, x_pos, _, parent):
        batch_size = z_hat.shape[0]
        logits = self.compute_logits(z_hat, x_pos, parent)  # [2B, 2B * world_size]
        hat_H_mlz, logs = self.compute_loss(logits)  # [2B]
        hat_H_mlz = (hat_H_mlz[:batch_size] + hat_H_mlz[batch_size:]) / 2
        return hat_H_mlz, logs, dict()

    def compute_logits(self, z_hat, x_pos, parent):
        z_pos_hat = x_pos if self.is_already_featurized else parent(x_pos, is_features=True, is_dist=True)
        zs = torch.cat([self.projector(z_hat), self.projector(z_pos_hat)], dim=0)
        n = zs.shape[0]
        is_dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        world = torch.distributed.get_world_size() if is_dist else 1
        logits = _LogitsTile.apply(zs, n)
        if world > 1:
            # every rank gathers the LOGIT TILES (not the embeddings), own tile first: the remote
            # columns are the other ranks' self-similarities and carry no gradient here
            tiles = gather_from_gpus(logits)
            rank = torch.distributed.get_rank()
            logits = torch.cat([tiles[rank]] + [t for r, t in enumerate(tiles) if r != rank], dim=1)
        return logits

    def compute_loss(self, logits):
        new_batch_size, n_cols = logits.shape
        n_classes = n_cols - 1  # the row's own similarity is masked out inside the kernel
        if self.effective_batch_size is not None:
            effective_n_classes = 2 * self.effective_batch_size - 1
            pos_add = -math.log((effective_n_classes - 1) / (n_classes - 1))  # added to the positive logit
        else:
            effective_n_classes, pos_add = n_classes, 0.0
        if self.is_train_temperature:
            inv_temp = torch.clamp(self.logit_scale.exp(), max=1 / self.temperature)
        else:
            inv_temp = torch.tensor(1 / self.temperature, device=logits.device)
        hat_H_m = math.log(effective_n_classes)
        hat_H_mlz = _InfoNCE.apply(logits.contiguous(), inv_temp, pos_add)
        logs = dict(I_q_zm=(hat_H_m - hat_H_mlz.mean()) / math.log(BASE_LOG), hat_H_m=hat_H_m / math.log(BASE_LOG),
                    n_negatives=n_classes)
        return hat_H_mlz, logs


class LossyZDistortion(nn.Module):
    """||z_hat - z + 1e-6||_1 per row against the encoder's own z (nn.PairwiseDistance(p=1))."""

    def __init__(self, p_norm=1):
        super().__init__()
        if p_norm != 1:
            raise NotImplementedError("the lossy-Z kernel implements the reference's p_norm=1 only")
        self.p_norm, self.eps = p_norm, 1e-6

    def forward(self, z_hat, _, p_Zlx, __):
        dist = _LossyZ.apply(z_hat, p_Zlx.base_dist.mean, self.eps)
        return dist, dict(), dict()


def prediction_loss(Y_hat, y, is_classification=True, agg_over_tasks="mean"):
    """lossyless/helpers.py:439-489: log loss (classification) or squared error, averaged over the target
    dimension (dim 1), then aggregated over tasks (trailing dimensions) -> [batch, 1] (or [batch, n_tasks])."""
    if is_classification:
        loss = F.cross_entropy(Y_hat, y.long(), reduction="none")
    else:
        loss = F.mse_loss(Y_hat, y, reduction="none")
    while loss.dim() < 3:  # atleast_ndim(loss, 3): n_tasks = 1 if not given
        loss = loss.unsqueeze(-1)
    batch_size, Y_dim = loss.shape[:2]
    loss = loss.view(batch_size, Y_dim, -1).mean(dim=1)  # [batch, n_tasks]
    if agg_over_tasks is None:
        return loss
    if agg_over_tasks in ("max", "min", "median"):
        return getattr(loss, agg_over_tasks)(keepdim=True, dim=1)[0]
    if agg_over_tasks in ("mean", "sum", "std"):
        return getattr(loss, agg_over_tasks)(keepdim=True, dim=1)
    raise ValueError(f"Unkown agg_over_tasks={agg_over_tasks}.")


class DirectDistortion(nn.Module):
    """Direct variational bound for vector targets (``data_mode="distribution"``, lossyless/distortions.py:41-172):
    decoder MLP + ``prediction_loss`` (squared error averaged over the target dimension for regression targets, log
    loss for classification targets), summed over what is left per sample.  Plain torch -- only the CPU plumbing
    config (banana) uses it; image decoders (cnn / balle, un-normalisation) are outside the hot path."""

    def __init__(self, z_dim, y_shape, arch=None, arch_kwargs=dict(), dataset=None, is_normalized=False,
                 data_mode="distribution", name=None, **kwargs):
        super().__init__()
        if data_mode != "distribution":
            raise NotImplementedError("image decoders (cnn / balle) are outside the B200 hot path")
        if is_normalized:
            raise NotImplementedError("Can curently only deal with normalized data if it's an image.")
        self.dataset, self.is_img_out, self.is_normalized = dataset, False, False
        self.q_YlZ = get_Architecture(arch or "mlp", **arch_kwargs)(z_dim, y_shape)
        self.kwargs = dict(kwargs)  # -> prediction_loss (is_classification, agg_over_tasks)
        self.kwargs.setdefault("is_classification", False)  # the vector configs (banana) regress their target
        self.reset_parameters()

    def reset_parameters(self):
        weights_init(self)

    def forward(self, z_hat, aux_target, _, __):
        Y_hat = self.q_YlZ(z_hat)
        neg_log_q_ylz = prediction_loss(Y_hat, aux_target, **self.kwargs).flatten(1).sum(-1)
        logs = dict(H_q_TlZ=neg_log_q_ylz.mean() / math.log(BASE_LOG))
        return neg_log_q_ylz, logs, dict(Y_hat=Y_hat[0].detach().cpu(), Y=aux_target[0].detach().cpu())
