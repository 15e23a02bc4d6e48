"""torch.hub entry points, same names / signatures / return values as the reference's
hubconf.py:22-52.  The reference downloads ``beta{beta:0.0e}_factorized_rate.pt`` from its
v1.0 GitHub release (hubconf.py:15); the same three files ship inside this package
(lossyless_b200/hub/weights/) because there is no network where this runs, and the URL is
only used as a fall-back."""
dependencies = ["torch", "numpy"]  # dependencies required for loading a model

import os

import torch

from hub import ClipCompressor as _ClipCompressor

PATH = "https://github.com/YannDubs/lossyless/releases/download/v1.0/beta{beta:0.0e}_factorized_rate.pt"
LOCAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lossyless_b200", "hub", "weights",
                     "beta{beta:0.0e}_factorized_rate.pt")
DEVICE = "cuda" if torch.cuda.is_available() else "cpu"


def _load_state_dict(beta):
    local = LOCAL.format(beta=beta)
    if os.path.exists(local):
        return torch.load(local, map_location="cpu", weights_only=True)
    return torch.hub.load_state_dict_from_url(PATH.format(beta=beta), progress=False)


def _entry(beta, device, **kwargs):
    compressor = _ClipCompressor(pretrained_state_dict=_load_state_dict(beta), device=device, **kwargs)
    return compressor, compressor.preprocess


def clip_compressor_b005(device=DEVICE, **kwargs):
    return _entry(0.05, device, **kwargs)


def clip_compressor_b001(device=DEVICE, **kwargs):
    return _entry(0.01, device, **kwargs)


def clip_compressor_b01(device=DEVICE, **kwargs):
    return _entry(0.1, device, **kwargs)


DOCSTRING = """
    Load the invariant CLIP compressor trained with beta={beta:.0e} (larger beta = stronger compression).

    Parameters
    ----------
    device : str
        Device on which to load the model.
    kwargs :
        Forwarded to ``ClipCompressor`` (e.g. ``clip_state_dict=`` for the CLIP ViT-B/32 weights).

    Return
    ------
    compressor : nn.Module
        ``compressor(X)`` on a batch of images returns decompressed representations,
        ``compressor.compress(X)`` the compressed byte strings; ``compress_dataset(dataset, file)`` /
        ``decompress_dataset(file)`` work on whole datasets.
    transform : callable
        Resize to (3,224,224) + CLIP normalisation, usable as a torchvision transform.
    """

clip_compressor_b005.__doc__ = DOCSTRING.format(beta=0.05)
clip_compressor_b001.__doc__ = DOCSTRING.format(beta=0.01)
clip_compressor_b01.__doc__ = DOCSTRING.format(beta=0.1)
