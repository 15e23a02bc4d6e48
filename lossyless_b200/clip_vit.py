"""CLIP ViT-B/32 image encoder on the B200 kernels.

Host-side mirror of ``clip.model.VisionTransformer`` (clip==1.0) as the reference obtains it
from ``clip.load("ViT-B/32")`` -> ``model.visual`` (hub/compressor.py:39-40,
lossyless/architectures.py:367-376): same constructor, same state-dict key names
(``conv1.weight``, ``class_embedding``, ``positional_embedding``, ``ln_pre.*``,
``transformer.resblocks.{i}.{attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.*, ln_1.*,
mlp.c_fc.*, mlp.c_proj.*, ln_2.*}``, ``ln_post.*``, ``proj``), so a CLIP checkpoint's ``visual.*``
tensors load with ``load_state_dict``.  The forward pass is one native call
(``ll_vit_forward``: tcgen05 GEMMs + fused kernels); the torch modules below only hold
parameters.  Inference only (the reference's shipped CLIP recipe freezes the encoder,
bin/clip/clip_hub.sh / config/finetune/freezer.yaml).  No CPU path.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _native as N

__all__ = ["VisionTransformer", "clip_preprocess", "CLIP_MEAN", "CLIP_STD", "load_visual_state_dict"]

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class _Linear(nn.Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n_out, n_in))
        self.bias = nn.Parameter(torch.zeros(n_out))


class _LayerNorm(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))


class _Attention(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = _Linear(width, width)


class _Mlp(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.c_fc = _Linear(width, 4 * width)
        self.c_proj = _Linear(4 * width, width)


class _ResBlock(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.attn = _Attention(width)
        self.ln_1 = _LayerNorm(width)
        self.mlp = _Mlp(width)
        self.ln_2 = _LayerNorm(width)


class _Transformer(nn.Module):
    def __init__(self, width, layers):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(width) for _ in range(layers)])


class _Conv(nn.Module):
    def __init__(self, width, patch):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(width, 3, patch, patch))


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution=224, patch_size=32, width=768, layers=12, heads=12, output_dim=512):
        super().__init__()
        self.input_resolution, self.patch_size = input_resolution, patch_size
        self.width, self.layers, self.heads, self.output_dim = width, layers, heads, output_dim
        self.conv1 = _Conv(width, patch_size)
        self.class_embedding = nn.Parameter(torch.empty(width))
        self.positional_embedding = nn.Parameter(torch.empty((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = _LayerNorm(width)
        self.transformer = _Transformer(width, layers)
        self.ln_post = _LayerNorm(width)
        self.proj = nn.Parameter(torch.empty(width, output_dim))
        self._native = None
        self.initialize_parameters()

    @torch.no_grad()
    def initialize_parameters(self, seed=0):
        """Deterministic CLIP-style random init (``CLIP.initialize_parameters`` scales); used when
        no pretrained weights are available (there is no network in this environment)."""
        g = torch.Generator().manual_seed(seed)
        W, L = self.width, self.layers
        proj_std = (W ** -0.5) * ((2 * L) ** -0.5)
        attn_std, fc_std = W ** -0.5, (2 * W) ** -0.5
        # every tensor comes from the seeded generator (never the global RNG): all ranks of a
        # multi-GPU job must hold the same encoder
        self.class_embedding.copy_(torch.randn(W, generator=g) * W ** -0.5)
        self.positional_embedding.copy_(torch.randn(self.positional_embedding.shape, generator=g) * W ** -0.5)
        self.proj.copy_(torch.randn(W, self.output_dim, generator=g) * W ** -0.5)
        self.conv1.weight.copy_(torch.randn(self.conv1.weight.shape, generator=g) * (3 * self.patch_size ** 2) ** -0.5)
        for blk in self.transformer.resblocks:
            blk.attn.in_proj_weight.copy_(torch.randn(3 * W, W, generator=g) * attn_std)
            blk.attn.out_proj.weight.copy_(torch.randn(W, W, generator=g) * proj_std)
            blk.mlp.c_fc.weight.copy_(torch.randn(4 * W, W, generator=g) * fc_std)
            blk.mlp.c_proj.weight.copy_(torch.randn(W, 4 * W, generator=g) * proj_std)
        self._native = None

    # -- native handle ------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        self._native = None  # device / dtype changes invalidate the kernel-side views
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._native = None
        return super().load_state_dict(*a, **k)

    def refresh_native_(self):
        """Drop the kernel-side weight views (call after modifying parameters in place)."""
        self._native = None

    def _build_native(self):
        dev = self.proj.device
        if dev.type != "cuda":
            raise N.NativeError("the CLIP ViT kernels need the module on a B200 (cuda); there is no CPU path")
        N.check(N.load().ll_check_device(dev.index if dev.index is not None else torch.cuda.current_device()), "ll_check_device")
        dt = self.proj.dtype
        if dt == torch.float32:
            dt = torch.float16  # reference: clip.load(...) on cuda holds fp16 weights (hub/compressor.py:39)
        keep = []

        def w16(t):
            t = t.detach().to(dt).contiguous()
            keep.append(t)
            return t.data_ptr()

        def f32(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        blocks = list(self.transformer.resblocks)
        arr = lambda vals: (ctypes.c_void_p * len(vals))(*vals)
        arrays = dict(
            in_proj_w=arr([w16(b.attn.in_proj_weight) for b in blocks]),
            in_proj_b=arr([f32(b.attn.in_proj_bias) for b in blocks]),
            out_proj_w=arr([w16(b.attn.out_proj.weight) for b in blocks]),
            out_proj_b=arr([f32(b.attn.out_proj.bias) for b in blocks]),
            ln1_w=arr([f32(b.ln_1.weight) for b in blocks]), ln1_b=arr([f32(b.ln_1.bias) for b in blocks]),
            ln2_w=arr([f32(b.ln_2.weight) for b in blocks]), ln2_b=arr([f32(b.ln_2.bias) for b in blocks]),
            fc_w=arr([w16(b.mlp.c_fc.weight) for b in blocks]), fc_b=arr([f32(b.mlp.c_fc.bias) for b in blocks]),
            cproj_w=arr([w16(b.mlp.c_proj.weight) for b in blocks]), cproj_b=arr([f32(b.mlp.c_proj.bias) for b in blocks]),
        )
        weights = N.VitWeights(
            conv1_w=w16(self.conv1.weight.reshape(self.width, -1)), class_emb=w16(self.class_embedding),
            pos_emb=w16(self.positional_embedding),
            ln_pre_w=f32(self.ln_pre.weight), ln_pre_b=f32(self.ln_pre.bias),
            ln_post_w=f32(self.ln_post.weight), ln_post_b=f32(self.ln_post.bias),
            proj_t=w16(self.proj.t()),
            **{k: ctypes.cast(v, ctypes.POINTER(ctypes.c_void_p)) for k, v in arrays.items()})
        cfg = N.VitConfig(width=self.width, layers=self.layers, heads=self.heads, patch=self.patch_size,
                          res=self.input_resolution, out_dim=self.output_dim, dtype=N._DTYPES[dt])
        handle = ctypes.c_void_p()
        N.call("ll_vit_create", ctypes.byref(cfg), ctypes.byref(weights), ctypes.byref(handle))
        self._native = dict(handle=handle, keep=keep, arrays=arrays, dtype=dt, ws=None, device=dev)
        return self._native

    def __del__(self):
        nat = getattr(self, "_native", None)
        if nat is not None:
            try:
                N.load().ll_vit_destroy(nat["handle"])
            except Exception:
                pass

    def set_chunk(self, images_per_pass):
        nat = self._native or self._build_native()
        N.call("ll_vit_set_chunk", nat["handle"], int(images_per_pass))
        nat["ws"] = None

    def set_fuse_layernorm(self, on):
        """ln_1 / ln_2 folded into the neighbouring GEMMs (default) or run as separate kernels."""
        nat = self._native or self._build_native()
        N.call("ll_vit_set_fuse_layernorm", nat["handle"], int(bool(on)))

    def launches(self, batch):
        nat = self._native or self._build_native()
        return int(N.load().ll_vit_launches(nat["handle"], int(batch)))

    def profile_begin(self):
        nat = self._native or self._build_native()
        N.call("ll_vit_profile_begin", nat["handle"])

    def profile_end(self):
        """-> (gemm_ms, gemm_flops, gemm_launches) accumulated since profile_begin()."""
        nat = self._native
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        N.call("ll_vit_profile_end", nat["handle"], ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
        return ms.value, fl.value, n.value

    def profile_breakdown(self):
        """-> dict of event-timed ms per kernel family for the last profile window."""
        arr = (ctypes.c_double * 4)()
        N.call("ll_vit_profile_breakdown", self._native["handle"], arr)
        return dict(gemm=arr[0], attention=arr[1], layernorm=arr[2], patchify_assemble=arr[3])

    @torch.no_grad()
    def forward(self, x, out=None):
        """x [B,3,R,R] (fp16 / bf16 / fp32, cuda, CLIP-normalised) -> z [B, output_dim] in the
        model's 16-bit dtype (the reference's fp16 ``self.clip(X)``).  ``out`` optionally receives
        the result (a contiguous [B, output_dim] view of a larger buffer, for pipelined callers)."""
        N.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.input_resolution or x.shape[3] != self.input_resolution:
            raise ValueError(f"expected [B,3,{self.input_resolution},{self.input_resolution}] images, got {tuple(x.shape)}")
        nat = self._native or self._build_native()
        x = x.contiguous()
        B = x.shape[0]
        if out is not None:
            if out.shape != (B, self.output_dim) or out.dtype != nat["dtype"] or not out.is_contiguous():
                raise ValueError("out must be a contiguous [B, output_dim] tensor of the model dtype")
            z = out
        else:
            z = torch.empty((B, self.output_dim), dtype=nat["dtype"], device=x.device)
        if B == 0:
            return z
        need = int(N.load().ll_vit_workspace_bytes(nat["handle"], B))
        if nat["ws"] is None or nat["ws"].numel() < need:
            nat["ws"] = torch.empty(need, dtype=torch.uint8, device=x.device)
        N.call("ll_vit_forward", nat["handle"], N.ptr(x), N.dtype_code(x), B, N.ptr(z), N.ptr(nat["ws"]),
               nat["ws"].numel(), N.stream())
        return z


def load_visual_state_dict(sd):
    """Accept a full CLIP state dict (``visual.*`` keys) or a visual-only one."""
    if any(k.startswith("visual.") for k in sd):
        sd = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
    return sd


def _convert_image_to_rgb(image):
    return image.convert("RGB") if hasattr(image, "convert") else image


def clip_preprocess(n_px=224):
    """``clip.clip._transform``: bicubic resize, centre crop, RGB, tensor, CLIP normalisation
    (lossyless/helpers.py:252,260 hold the same constants)."""
    from torchvision import transforms as T

    return T.Compose([
        T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC),
        T.CenterCrop(n_px),
        _convert_image_to_rgb,
        T.ToTensor(),
        T.Normalize(CLIP_MEAN, CLIP_STD),
    ])
