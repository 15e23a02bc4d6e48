"""CLIP preprocessing on the GPU: the transform ``clip.load`` hands back (``clip.clip._transform`` =
torchvision ``Resize(n_px, BICUBIC)`` -> ``CenterCrop(n_px)`` -> RGB -> ``ToTensor`` -> ``Normalize``), which
the reference runs in CPU DataLoader workers before ``compress_dataset`` (hub/compressor.py:39,186;
utils/data/images.py:383-389).  Here the uint8 pixels go to the device as they are (6-16x fewer H2D bytes
than the fp16 224x224 tensor) and one native call does the rest.

The resize is Pillow's 8-bit ``ImagingResample`` restated exactly (bit-identical to PIL, see
tests/test_preprocess_gpu.py): this module computes Pillow's per-output-pixel windows and 22-bit
fixed-point bicubic coefficients in double precision with Pillow's own formulas; the kernels
(csrc/preprocess.cu) do the integer accumulation, clamping, normalisation and the fp16 cast.
"""
import ctypes
import math

import numpy as np
import torch

from . import _native as N
from .clip_vit import CLIP_MEAN, CLIP_STD

__all__ = ["GpuClipPreprocess", "resized_size", "resample_tables"]

_PRECISION_BITS = 32 - 8 - 2


def _bicubic(x, a=-0.5):
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resized_size(h, w, size):
    """torchvision ``Resize(int)``: smaller edge -> ``size``, the other ``int(size * long / short)``."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def resample_tables(in_size, out_size, first, count):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for output indices [first, first+count):
    (coef int32 [count, ksize], bounds int32 [count, 2])."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale  # bicubic support
    ksize = int(math.ceil(support)) * 2 + 1
    coef = np.zeros((count, ksize), dtype=np.int32)
    bounds = np.zeros((count, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for j in range(count):
        xx = first + j
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in k:
            ww += v
        if ww != 0.0:
            k = [v / ww for v in k]
        for x, v in enumerate(k):
            coef[j, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[j] = (xmin, xmax)
    return coef, bounds


class GpuClipPreprocess:
    """``preprocess(images_uint8)``: [B,H,W,3] (or [B,3,H,W] with ``chw=True``) uint8 on the GPU ->
    [B,3,n_px,n_px] CLIP-normalised, ``dtype`` fp16 (what ``compress_dataset`` feeds the encoder) or fp32
    (what the CPU transform returns).  All images of a batch share one size."""

    def __init__(self, n_px=224, mean=CLIP_MEAN, std=CLIP_STD):
        self.n_px = n_px
        # torchvision Normalize: mean / std as float32 tensors
        self.mean = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in mean])
        self.std = (ctypes.c_float * 3)(*[float(np.float32(s)) for s in std])
        self._tables = {}

    def _get_tables(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._tables:
            n = self.n_px
            new_h, new_w = resized_size(h, w, n)
            top = int(round((new_h - n) / 2.0))  # torchvision center_crop
            left = int(round((new_w - n) / 2.0))
            ch, bh = resample_tables(w, new_w, left, n)
            cv, bv = resample_tables(h, new_h, top, n)
            to = lambda a: torch.from_numpy(a).to(device)
            self._tables[key] = (to(ch), to(bh), ch.shape[1], to(cv), to(bv), cv.shape[1])
        return self._tables[key]

    def __call__(self, images, chw=False, dtype=torch.float16):
        N.require_cuda(images, "images")
        if images.dtype != torch.uint8 or images.dim() != 4:
            raise ValueError("expected a uint8 [B,H,W,3] (or [B,3,H,W] with chw=True) tensor")
        images = images.contiguous()
        if chw:
            B, C, H, W = images.shape
        else:
            B, H, W, C = images.shape
        if C != 3:
            raise ValueError("expected RGB images (3 channels)")
        ch, bh, kh, cv, bv, kv = self._get_tables(H, W, images.device)
        n = self.n_px
        tmp = torch.empty((B, H, n, 3), dtype=torch.uint8, device=images.device)
        out = torch.empty((B, 3, n, n), dtype=dtype, device=images.device)
        N.call("ll_clip_preprocess", N.ptr(images), B, H, W, 1 if chw else 0, N.ptr(ch), N.ptr(bh), kh, N.ptr(cv), N.ptr(bv), kv,
               self.mean, self.std, n, N.ptr(tmp), N.ptr(out), N.dtype_code(out), N.stream())
        return out
