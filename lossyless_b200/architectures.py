"""Architectures on the hot path (reference lossyless/architectures.py): ``get_Architecture``
(:42-85), ``MLP / FlattenMLP / FlattenLinear`` (:94-232, the projector and the banana plumbing
config) and ``PretrainedSSL`` for ``clip_vit`` (:312-379) backed by the B200 ViT.  Resnet / CNN /
BALLE / SimCLR / SwAV / CLIP-RN50 are other encoders outside SURVEY.md section 8 and raise."""
from functools import partial

import torch
import torch.nn as nn

from .clip_vit import VisionTransformer, load_visual_state_dict
from .helpers import batch_flatten, batch_unflatten, prod, weights_init

__all__ = ["get_Architecture", "MLP", "FlattenMLP", "FlattenLinear", "PretrainedSSL"]

_OUT_OF_SCOPE = ("resnet", "cnn", "balle", "clip_rn50", "simclr", "swav")


def get_Architecture(mode, **kwargs):
    """Uninstantiated architecture: ``Architecture(in_shape, out_shape)``."""
    if mode == "mlp":
        return partial(FlattenMLP, **kwargs)
    if mode == "linear":
        return partial(FlattenLinear, **kwargs)
    if mode == "identity":
        return torch.nn.Identity
    if mode == "clip":
        return partial(PretrainedSSL, model="clip_vit", **kwargs)
    if mode in _OUT_OF_SCOPE:
        raise NotImplementedError(f"architecture {mode!r} is outside the B200 hot path (SURVEY.md section 8); "
                                  "available: mlp, linear, identity, clip")
    raise ValueError(f"Unkown mode={mode}.")


def _activation(name):
    return getattr(nn, name)


def _norm(name):
    if name in (None, "identity"):
        return nn.Identity
    if name == "batch":
        return nn.BatchNorm1d
    raise ValueError(f"Unkown normalization={name}.")


class MLP(nn.Module):
    """in -> [hid, norm, act, dropout] x n_hid_layers -> out (no bias before a batch norm)."""

    def __init__(self, in_dim, out_dim, n_hid_layers=1, hid_dim=128, norm_layer="identity", activation="ReLU", dropout_p=0):
        super().__init__()
        self.in_dim, self.out_dim, self.n_hid_layers, self.hid_dim = in_dim, out_dim, n_hid_layers, hid_dim
        Act, Norm = _activation(activation), _norm(norm_layer)
        Drop = nn.Dropout if dropout_p > 0 else nn.Identity
        is_bias = Norm == nn.Identity
        layers, d = [], in_dim
        for _ in range(n_hid_layers):
            layers += [nn.Linear(d, hid_dim, bias=is_bias), Norm(hid_dim), Act(), Drop(p=dropout_p) if dropout_p > 0 else Drop()]
            d = hid_dim
        layers += [nn.Linear(hid_dim, out_dim)]
        self.module = nn.Sequential(*layers)
        self.reset_parameters()

    def forward(self, X):
        X, shape = batch_flatten(X)
        return batch_unflatten(self.module(X), shape)

    def reset_parameters(self):
        weights_init(self)


class FlattenMLP(MLP):
    def __init__(self, in_shape, out_shape, **kwargs):
        self.in_shape = [in_shape] if isinstance(in_shape, int) else list(in_shape)
        self.out_shape = [out_shape] if isinstance(out_shape, int) else list(out_shape)
        super().__init__(in_dim=prod(self.in_shape), out_dim=prod(self.out_shape), **kwargs)

    def forward(self, X):
        X = X.flatten(start_dim=X.ndim - len(self.in_shape))
        return super().forward(X).unflatten(dim=-1, sizes=self.out_shape)


class FlattenLinear(nn.Linear):
    def __init__(self, in_shape, out_shape, **kwargs):
        self.in_shape = [in_shape] if isinstance(in_shape, int) else list(in_shape)
        self.out_shape = [out_shape] if isinstance(out_shape, int) else list(out_shape)
        super().__init__(in_features=prod(self.in_shape), out_features=prod(self.out_shape), **kwargs)

    def forward(self, X):
        X = X.flatten(start_dim=X.ndim - len(self.in_shape))
        return super().forward(X).unflatten(dim=-1, sizes=self.out_shape)


class PretrainedSSL(nn.Module):
    """Pretrained SSL encoder; only ``model="clip_vit"`` (CLIP ViT-B/32, 3x224x224 -> 512).

    The reference calls ``clip.load`` (a download) and ``.float()``s the tower for AMP training;
    here the weights come from ``clip_state_dict`` (dict or path; default: seeded random init).
    ``is_frozen=True`` (default; the reference's shipped CLIP recipe, bin/clip/clip_hub.sh +
    config/finetune/freezer.yaml): the tower runs in fp16 through ``ll_vit_forward`` and the float32
    output carries no gradient.  ``is_frozen=False`` (the reference's ``PretrainedSSL`` without the
    ``Freezer`` callback, lossyless/architectures.py:351-376): float32 master parameters, 16-bit
    ``compute_dtype`` kernels, and a backward pass through the whole tower (``clip_vit._VitTrainFn``)."""

    def __init__(self, in_shape, out_shape, model="clip_vit", clip_state_dict=None, is_frozen=True,
                 compute_dtype=torch.bfloat16):
        super().__init__()
        if model != "clip_vit":
            raise NotImplementedError(f"PretrainedSSL model {model!r} is outside the B200 hot path; only clip_vit")
        self.in_shape = in_shape
        self.out_shape = [out_shape] if isinstance(out_shape, int) else list(out_shape)
        self.model, self.out_dim = model, prod(self.out_shape)
        assert self.out_dim == 512
        assert self.in_shape[0] == 3 and self.in_shape[1] == self.in_shape[2] == 224
        self.encoder = VisionTransformer(224, 32, 768, 12, 12, 512)
        if clip_state_dict is not None:
            if not isinstance(clip_state_dict, dict):
                clip_state_dict = torch.load(clip_state_dict, map_location="cpu", weights_only=True)
            self.encoder.load_state_dict(load_visual_state_dict(clip_state_dict))
        self.is_frozen = is_frozen
        if is_frozen:
            self.encoder.half()
            for p in self.encoder.parameters():
                p.requires_grad_(False)
        else:
            self.encoder.unfreeze_(compute_dtype)

    def forward(self, X):
        z = self.encoder(X).float()
        return z.unflatten(dim=-1, sizes=self.out_shape)

    def forward_views(self, Xs):
        """``forward(torch.cat(Xs))`` without the concatenated copy (CondDist.forward_many: the two views of a step)."""
        if not hasattr(self.encoder, "forward_views"):
            return self.forward(torch.cat(list(Xs), dim=0))
        z = self.encoder.forward_views(Xs).float()
        return z.unflatten(dim=-1, sizes=self.out_shape)

    def reset_parameters(self):
        pass  # pretrained / seeded weights are kept
