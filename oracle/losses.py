"""Oracle restatement of the reference's distortion / loss arithmetic (plain torch fp32 CPU).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).

Follows, function by function:
  * lossyless/distortions.py:305-338  ContrastiveDistortion.compute_logits
  * lossyless/distortions.py:340-405  ContrastiveDistortion.compute_loss
  * lossyless/distortions.py:266-303  ContrastiveDistortion.forward (row pairing)
  * lossyless/helpers.py:873-893      GatherFromGpus (forward concat order; backward keeps
                                      only the local gradient)
  * lossyless/distortions.py:408-450  LossyZDistortion (nn.PairwiseDistance(p=1), eps=1e-6
                                      added to the difference)
  * lossyless/learnable_compressors.py:241-275  LearnableCompressor.loss
  * lossyless/helpers.py:768-870      Annealer
"""
import math

import torch
import torch.nn.functional as F

BASE_LOG = 2  # lossyless/helpers.py:27


def projector(proj, z):
    """FlattenMLP(512 -> 128 -> ReLU -> 256), lossyless/architectures.py:94-200 with
    config/distortion/base_contrastive.yaml:13-16.  proj = dict(w1,b1,w2,b2)."""
    h = F.relu(z @ proj["w1"].t() + proj["b1"])
    return h @ proj["w2"].t() + proj["b2"]


def contrastive_local_logits(z_hat, z_pos_hat, proj=None, is_cosine=True):
    """zs @ zs.T for one rank: [2B,2B]."""
    z = projector(proj, z_hat) if proj is not None else z_hat
    zp = projector(proj, z_pos_hat) if proj is not None else z_pos_hat
    zs = torch.cat([z, zp], dim=0)
    if is_cosine:
        zs = F.normalize(zs, dim=1, p=2)
    return zs @ zs.t()


def contrastive_gather(tiles, rank):
    """The multi-rank quirk: every rank all_gathers the [2B,2B] *logit tiles* and puts its
    own first (distortions.py:327-336); remote columns are other ranks' self-similarities."""
    others = tiles[:rank] + tiles[rank + 1:]
    return torch.cat([tiles[rank]] + others, dim=1) if others else tiles[rank]


def contrastive_loss(logits, temperature=0.01, logit_scale=None, effective_batch_size=None):
    """compute_loss on [2B, 2B*W] logits -> (hat_H_mlz [2B], logs)."""
    n_classes = logits.size(1)
    nb = logits.size(0)
    B = nb // 2
    mask = ~torch.eye(nb).bool()
    mask = torch.cat([mask, torch.ones(nb, n_classes - nb).bool()], dim=1)
    n_classes -= 1
    logits = logits[mask].view(nb, n_classes).clone()
    ar = torch.arange(B)
    pos_idx = torch.cat([ar + B - 1, ar], dim=0)
    if effective_batch_size is not None:
        effective_n_classes = 2 * effective_batch_size - 1
        to_mult = (effective_n_classes - 1) / (n_classes - 1)
        to_add = -math.log(to_mult) * torch.ones_like(logits[:, 0:1])
        logits = logits.scatter_add(1, pos_idx.unsqueeze(1), to_add)
    else:
        effective_n_classes = n_classes
    if logit_scale is not None:
        temp = 1 / torch.clamp(logit_scale.exp(), max=1 / temperature)
    else:
        temp = temperature
    logits = logits / temp
    hat_H_m = math.log(effective_n_classes)
    hat_H_mlz = F.cross_entropy(logits, pos_idx, reduction="none")
    logs = dict(I_q_zm=(hat_H_m - hat_H_mlz.mean()) / math.log(BASE_LOG),
                hat_H_m=hat_H_m / math.log(BASE_LOG), n_negatives=n_classes)
    return hat_H_mlz, logs


def contrastive_distortion(z_hat, z_pos_hat, proj, rank=0, other_tiles=None, **kw):
    """Full ContrastiveDistortion.forward for one rank -> distortions [B]."""
    B = z_hat.shape[0]
    tile = contrastive_local_logits(z_hat, z_pos_hat, proj)
    if other_tiles is not None:
        tiles = list(other_tiles)
        tiles.insert(rank, tile)
        logits = contrastive_gather(tiles, rank)
    else:
        logits = tile
    h, logs = contrastive_loss(logits, **kw)
    return (h[:B] + h[B:]) / 2, logs


def lossy_z_distortion(z_hat, z, p=1, eps=1e-6):
    """nn.PairwiseDistance(p=1): ||z_hat - z + eps||_p per row."""
    return torch.norm(z_hat - z + eps, p=p, dim=-1)


class Annealer:
    """lossyless/helpers.py:768-870 (call-with-n_update_calls form used by .loss)."""

    def __init__(self, initial_value, final_value, n_steps_anneal, start_step=0, default=None, mode="geometric"):
        if n_steps_anneal < 0:
            n_steps_anneal *= -1
            initial_value, final_value = final_value, initial_value
        self.i, self.f, self.n, self.s = initial_value, final_value, n_steps_anneal, start_step
        self.default = default if default is not None else initial_value
        self.mode = mode.lower()
        if self.mode == "linear":
            self.factor = (self.f - self.i) / self.n
        elif self.mode == "geometric":
            self.factor = (self.f / self.i) ** (1 / self.n)

    def __call__(self, n_update_calls):
        if self.s > n_update_calls:
            return self.default
        k = n_update_calls - self.s
        if self.mode != "constant" and n_update_calls < self.n + self.s:
            return self.i * self.factor ** k if self.mode == "geometric" else self.i + self.factor * k
        return self.f


def rd_loss(rates, distortions, curr_beta, final_beta, labda):
    """LearnableCompressor.loss: value uses final_beta, gradient uses curr_beta."""
    rates, distortions = rates.float(), distortions.float()
    rate, distortion = rates.mean(0), distortions.mean(0)
    beta_rate = curr_beta * rate
    beta_rate = beta_rate - beta_rate.detach() + final_beta * rate.detach()
    loss = labda * distortion + beta_rate
    logs = dict(loss=loss / math.log(BASE_LOG), rate=rate / math.log(BASE_LOG),
                distortion=distortion / math.log(BASE_LOG), beta=curr_beta,
                loose_loss=(labda * distortions + final_beta * rates).mean().detach() / math.log(BASE_LOG))
    return loss, logs
