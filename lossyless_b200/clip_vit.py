"""CLIP ViT-B/32 image encoder on the B200 kernels.

Host-side mirror of ``clip.model.VisionTransformer`` (clip==1.0) as the reference obtains it
from ``clip.load("ViT-B/32")`` -> ``model.visual`` (hub/compressor.py:39-40,
lossyless/architectures.py:367-376): same constructor, same state-dict key names
(``conv1.weight``, ``class_embedding``, ``positional_embedding``, ``ln_pre.*``,
``transformer.resblocks.{i}.{attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.*, ln_1.*,
mlp.c_fc.*, mlp.c_proj.*, ln_2.*}``, ``ln_post.*``, ``proj``), so a CLIP checkpoint's ``visual.*``
tensors load with ``load_state_dict``.  The forward pass is one native call
(``ll_vit_forward``: tcgen05 GEMMs + fused kernels); the torch modules below only hold
parameters.  Two execution paths, no CPU path:

* frozen / inference (default; the reference's shipped CLIP recipe freezes the encoder,
  bin/clip/clip_hub.sh / config/finetune/freezer.yaml): one ``ll_vit_forward`` call per batch;
* fine-tuning (``VisionTransformer.unfreeze_()``; the reference trains through
  ``clip.load(...).visual.float()`` with torch autograd under AMP, lossyless/architectures.py:351-376):
  float32 master parameters, 16-bit (bf16 by default) compute copies, a layer-by-layer forward that keeps
  the activations, and a hand-written backward (``_VitTrainFn``) whose contractions all run on the tcgen05
  GEMM (``ll_gemm`` on transposed copies) plus the kernels of csrc/vit_train.cu.  torch is used there for
  allocation, reshapes / slices (data movement) and the three tiny batch reductions of the embedding
  gradients only.
"""
import ctypes
import os

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
        self._trainable = False
        self._train_dtype = torch.bfloat16
        self._flat_enabled, self._flat = False, None
        self._train_cache = None
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
        self._flat = None    # ... and give every parameter its own storage again (the flat state is rebuilt lazily)
        self._train_cache = None
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

    def set_direct_patch(self, on):
        """Patch-embedding GEMM reads the image batch through a 5-D TMA map (default) or an im2col copy (A/B knob)."""
        nat = self._native or self._build_native()
        N.call("ll_vit_set_direct_patch", nat["handle"], int(bool(on)))

    def set_cls_last_block(self, on):
        """Last block's out-projection / ln_2 / MLP on the class-token rows only (default) or on every row (A/B knob)."""
        nat = self._native or self._build_native()
        N.call("ll_vit_set_cls_last_block", nat["handle"], int(bool(on)))

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

    # -- fine-tuning ---------------------------------------------------------------------
    def unfreeze_(self, compute_dtype=torch.bfloat16):
        """Make the encoder trainable: float32 master parameters (``clip.load(...).float()``,
        lossyless/architectures.py:376), ``requires_grad``, 16-bit compute in ``compute_dtype``
        (the AMP autocast dtype).  ``forward`` then records a backward pass whenever grad is enabled."""
        if compute_dtype not in (torch.bfloat16, torch.float16):
            raise ValueError("compute_dtype must be torch.bfloat16 or torch.float16")
        self.float()
        for p in self.parameters():
            p.requires_grad_(True)
        self._trainable, self._train_dtype, self._train_cache = True, compute_dtype, None
        import weakref

        ref = weakref.ref(self)
        for p in self.parameters():
            p._ll_vit = ref  # lets helpers.get_optimizer find the tower a parameter belongs to
        return self

    def enable_flat_training_(self, on=True):
        """Flat master / gradient / moment buffers and native optimiser kernels for this tower (``FlatTrainState``);
        switched on by ``helpers.FlatAdamW``.  In this mode the backward pass leaves the gradients in the flat buffers
        (``self.flat_state().grad_view(name)``) instead of returning them to autograd, so ``p.grad`` stays ``None``."""
        self._flat_enabled = bool(on)
        if not on:
            self._flat = None
        self._train_cache = None
        return self

    def flat_state(self):
        if not self._flat_enabled:
            return None
        if self._flat is None or self._flat.dt != self._train_dtype or self._flat.device != self.proj.device:
            self._flat = FlatTrainState(self, self._train_dtype)
            self._train_cache = None
        return self._flat

    def freeze_(self):
        for p in self.parameters():
            p.requires_grad_(False)
        self._trainable, self._train_cache = False, None
        self._flat_enabled, self._flat = False, None
        return self

    def forward_views(self, xs):
        """``forward(torch.cat(xs))`` for a list of image batches (the two views of a contrastive step) without
        materialising the concatenation: the fine-tuning path patchifies each view straight into its slice."""
        return self.forward(tuple(xs))

    def forward(self, x, out=None):
        """x [B,3,R,R] (fp16 / bf16 / fp32, cuda, CLIP-normalised) -> z [B, output_dim].
        Frozen encoder / no grad: the model's 16-bit dtype (the reference's fp16 ``self.clip(X)``);
        ``out`` optionally receives the result (a contiguous [B, output_dim] view of a larger buffer, for
        pipelined callers).  Unfrozen encoder with grad enabled: float32 z carrying the backward pass."""
        if self._trainable and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if out is not None:
                raise ValueError("out= is not supported on the fine-tuning path")
            names, params = zip(*self.named_parameters())
            return _VitTrainFn.apply(self, x, names, *params)
        if isinstance(x, (tuple, list)):
            x = torch.cat(list(x), dim=0)
        if self._native is not None and self._native.get("stale"):
            self._native = None  # parameters were updated by an optimiser step since the views were made
        with torch.no_grad():
            return self._forward_frozen(x, out)

    def _forward_frozen(self, x, out=None):
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


# ------------------------------------------------------------------------------ flat training state
def _is_matrix(name):
    """Parameters whose gradient is the output of a weight-gradient GEMM (16-bit, as autocast produces it)."""
    return name in ("proj", "conv1.weight") or name.endswith("in_proj_weight") or (
        name.endswith(".weight") and any(k in name for k in ("out_proj", "c_fc", "c_proj")))


class FlatTrainState:
    """Flat storage of an unfrozen tower's training state (``helpers.FlatAdamW`` switches it on).

    Every parameter becomes a view into one of two float32 master buffers -- ``mat`` (the GEMM weights, 85 M elements)
    and ``vec`` (biases, LayerNorm gains / biases, embeddings) -- laid out in the order the backward pass completes
    them (proj, block L-1 ... block 0, conv1), each tensor starting on a 16-byte boundary.  Alongside: the 16-bit
    compute copy of ``mat`` (what autocast would cast per call), the 16-bit gradient buffer the weight-gradient GEMMs
    write straight into, the float32 gradient buffer of ``vec``, and the AdamW moments.  What this buys over per-tensor
    torch state: one cast / norm / optimiser launch per buffer instead of ~10 per tensor (150 tensors), gradient
    all-reduces on contiguous slices with no packing, launched block by block from INSIDE the backward pass, and the
    next step's 16-bit weights written by the optimiser kernel itself."""

    ALIGN = 8  # elements: 16 bytes in the 16-bit buffers

    def __init__(self, vit, dt):
        params = dict(vit.named_parameters())
        dev = vit.proj.device
        order = ["proj", "ln_post.weight", "ln_post.bias"]
        self.block_first = {}
        for i in reversed(range(vit.layers)):
            p = f"transformer.resblocks.{i}."
            order += [p + "mlp.c_proj.weight", p + "mlp.c_proj.bias", p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", p + "ln_2.weight",
                      p + "ln_2.bias", p + "attn.out_proj.weight", p + "attn.out_proj.bias", p + "attn.in_proj_weight",
                      p + "attn.in_proj_bias", p + "ln_1.weight", p + "ln_1.bias"]
        order += ["ln_pre.weight", "ln_pre.bias", "positional_embedding", "class_embedding", "conv1.weight"]
        if set(order) != set(params):
            raise RuntimeError("FlatTrainState: unexpected parameter set")
        self.dt, self.device = dt, dev
        self.slots = {}  # name -> (group, offset, numel, shape)
        sizes = dict(mat=0, vec=0)
        for n in order:
            g = "mat" if _is_matrix(n) else "vec"
            numel = params[n].numel()
            self.slots[n] = (g, sizes[g], numel, tuple(params[n].shape))
            sizes[g] += (numel + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.names = order
        self.master = {g: torch.zeros(sizes[g], dtype=torch.float32, device=dev) for g in sizes}
        with torch.no_grad():
            for n in order:
                v = self.view(self.master, n)
                v.copy_(params[n].detach().float())
                params[n].data = v  # the module's parameters now alias the flat master
        self.w16 = torch.empty(sizes["mat"], dtype=dt, device=dev)
        self.grad = dict(mat=torch.zeros(sizes["mat"], dtype=dt, device=dev), vec=torch.zeros(sizes["vec"], dtype=torch.float32, device=dev))
        self.m = self.v = None  # AdamW moments, created by the first step
        self.step_count = 0
        self.cast_versions = None
        self.scratch = torch.empty(int(N.load().ll_sumsq_scratch_floats()), dtype=torch.float32, device=dev)
        # all-reduce slices of the matrix gradients in completion order: [proj + last block], one per block, conv1
        cuts = [0]
        for i in reversed(range(vit.layers)):
            g, off, numel, _ = self.slots[f"transformer.resblocks.{i}.attn.in_proj_weight"]
            cuts.append(off + (numel + self.ALIGN - 1) // self.ALIGN * self.ALIGN)
        cuts.append(sizes["mat"])
        self.mat_slices = [(a, b) for a, b in zip(cuts[:-1], cuts[1:])]  # L block slices (the first holds proj too) + conv1
        self.sync, self.sync_group, self.sync_world, self.pending = False, None, 1, []
        self.has_grads = False

    def view(self, bufs, name):
        g, off, numel, shape = self.slots[name]
        buf = bufs[g] if isinstance(bufs, dict) else bufs
        return buf[off: off + numel].view(shape)

    def grad_view(self, name):
        return self.view(self.grad, name)

    def refresh_w16(self, versions):
        """16-bit compute copy of the matrices: rewritten only if something other than ``adamw_step`` touched the
        parameters since it was last written."""
        if self.cast_versions != versions:
            N.call("ll_cast_flat", N.ptr(self.master["mat"]), N.ptr(self.w16), N._DTYPES[self.dt], self.w16.numel(), N.stream())
            self.cast_versions = versions

    # -- data-parallel gradient exchange, launched from inside the backward pass ------------------------
    def begin_backward(self):
        self.grad["vec"].zero_()  # the bias / LayerNorm kernels accumulate with atomics
        self.pending = []
        self.has_grads = True

    def reduce_mat_slice(self, k):
        """Block k (in completion order) of the matrix gradients is final: start its all-reduce (sum) while the
        backward pass of the earlier blocks keeps the tensor cores busy."""
        if self.sync:
            a, b = self.mat_slices[k]
            self.pending.append(torch.distributed.all_reduce(self.grad["mat"][a:b], group=self.sync_group, async_op=True))

    def reduce_vec(self):
        if self.sync:
            self.pending.append(torch.distributed.all_reduce(self.grad["vec"], group=self.sync_group, async_op=True))

    def wait_reductions(self):
        for w in self.pending:
            w.wait()
        self.pending = []

    # -- optimiser ----------------------------------------------------------------------------------------
    def sumsq(self, out):
        """out[0] = |grad|^2 over both buffers (before the 1 / world scaling)."""
        N.call("ll_sumsq", N.ptr(self.grad["mat"]), self.grad["mat"].numel(), N._DTYPES[self.dt], N.ptr(self.scratch), N.ptr(out), 0, N.stream())
        N.call("ll_sumsq", N.ptr(self.grad["vec"]), self.grad["vec"].numel(), N.LL_F32, N.ptr(self.scratch), N.ptr(out), 1, N.stream())
        return out

    def adamw_step(self, lr, betas, eps, weight_decay, grad_scale=None):
        if self.m is None:
            self.m = {g: torch.zeros_like(self.master[g]) for g in self.master}
            self.v = {g: torch.zeros_like(self.master[g]) for g in self.master}
        self.step_count += 1
        sc = N.ptr(grad_scale) if grad_scale is not None else None
        N.call("ll_adamw_flat", N.ptr(self.master["mat"]), N.ptr(self.grad["mat"]), N._DTYPES[self.dt], N.ptr(self.m["mat"]), N.ptr(self.v["mat"]),
               N.ptr(self.w16), N._DTYPES[self.dt], self.master["mat"].numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
               float(weight_decay), self.step_count, sc, N.stream())
        N.call("ll_adamw_flat", N.ptr(self.master["vec"]), N.ptr(self.grad["vec"]), N.LL_F32, N.ptr(self.m["vec"]), N.ptr(self.v["vec"]),
               None, N.LL_BF16, self.master["vec"].numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
               float(weight_decay), self.step_count, sc, N.stream())
        self.has_grads = False


# ------------------------------------------------------------------------------ fine-tuning path
def _gemm(A, W, bias=None, residual=None, act=0):
    """D[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual) on the tcgen05 GEMM; 16-bit operands."""
    M, K = A.shape
    Nn = W.shape[0]
    D = torch.empty((M, Nn), dtype=A.dtype, device=A.device)
    N.call("ll_gemm", N.ptr(A), N.ptr(W), N.ptr(bias), N.ptr(residual), N.ptr(D), M, Nn, K, N.dtype_code(A), act, N.stream())
    return D


def _ln(x, w, b):
    h = torch.empty_like(x)
    N.call("ll_layernorm", N.ptr(x), x.shape[0], x.shape[1], N.ptr(w), N.ptr(b), N.ptr(h), N.dtype_code(x), N.stream())
    return h


def _transpose_pad(t):
    """[R, C] -> [C, ceil64(R)] zero padded (R becomes the K of a weight-gradient GEMM)."""
    R, C = t.shape
    Rp = (R + 63) // 64 * 64
    out = torch.empty((C, Rp), dtype=t.dtype, device=t.device)
    N.call("ll_transpose_pad", N.ptr(t), N.ptr(out), R, C, Rp, N.stream())
    return out


# Weight gradients: "tn" = ll_gemm_tn, the tcgen05 GEMM reading both operands MN-major straight from the row-major
# activations / gradients (default); "transpose" = K-major GEMM on zero-padded transposed copies (the first
# implementation, kept for A/B runs and as a cross-check in the tests).
WGRAD_IMPL = "tn"


def _wgrad(gy, x, out=None):
    """dW [N, K] = gy[M, N]^T @ x[M, K] (16-bit result, as autocast produces it); ``out`` = where to put it (a view
    of the flat gradient buffer)."""
    if WGRAD_IMPL == "tn":
        D = out if out is not None else torch.empty((gy.shape[1], x.shape[1]), dtype=gy.dtype, device=gy.device)
        N.call("ll_gemm_tn", N.ptr(gy), N.ptr(x), N.ptr(D), gy.shape[0], gy.shape[1], x.shape[1], N.dtype_code(gy), N.stream())
        return D
    D = _gemm(_transpose_pad(gy), _transpose_pad(x))
    return D if out is None else out.copy_(D)


def _wgrad_bias(gy, x, out_w=None, out_b=None):
    """(dW, db): the weight gradient as above and db = column sums of gy (float32); on the transposing path the
    sums are taken while gy is being transposed so that it is read once.  ``out_w`` / ``out_b``: destinations in the
    flat gradient buffers (``out_b`` is accumulated into: it must be zero on entry)."""
    if WGRAD_IMPL == "tn" or out_w is not None:
        return _wgrad(gy, x, out_w), _colsum(gy, out_b)
    R, C = gy.shape
    Rp = (R + 63) // 64 * 64
    gyt = torch.empty((C, Rp), dtype=gy.dtype, device=gy.device)
    db = torch.zeros(C, dtype=torch.float32, device=gy.device)
    N.call("ll_transpose_pad_colsum", N.ptr(gy), N.ptr(gyt), R, C, Rp, N.dtype_code(gy), N.ptr(db), N.stream())
    return _gemm(gyt, _transpose_pad(x)), db


def _colsum(t, out=None):
    if out is None:
        out = torch.zeros(t.shape[1], dtype=torch.float32, device=t.device)
    N.call("ll_colsum", N.ptr(t), t.shape[0], t.shape[1], N.dtype_code(t), N.ptr(out), N.stream())
    return out


def _ln_backward(x, g_h, gamma, g_res=None, out_w=None, out_b=None):
    g_x = torch.empty_like(x)
    g_w = out_w if out_w is not None else torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
    g_b = out_b if out_b is not None else torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
    N.call("ll_layernorm_backward", N.ptr(x), N.ptr(g_h), N.ptr(gamma), N.ptr(g_res), N.ptr(g_x), N.ptr(g_w), N.ptr(g_b),
           x.shape[0], x.shape[1], N.dtype_code(x), N.stream())
    return g_x, g_w, g_b


# Native per-block runner (csrc/vit_train_runner.cu): one call per block and direction instead of ~13 / ~26 Python-issued
# launches.  False = the Python-sequenced path (kept as a cross-check in the tests and for A/B runs).
NATIVE_BLOCKS = os.environ.get("LL_NATIVE_BLOCKS", "1") != "0"
# The last block's out-projection / ln_2 / MLP (forward, data and weight gradients) on the class-token rows only -- the only
# rows VisionTransformer.forward keeps (x[:, 0, :]); native runner only.  False = every row (A/B, cross-check).
CLS_LAST_BLOCK = os.environ.get("LL_CLS_LAST_BLOCK", "1") != "0"

_BLOCK_FIELDS = (("in_proj_w", "attn.in_proj_weight"), ("out_proj_w", "attn.out_proj.weight"), ("fc_w", "mlp.c_fc.weight"),
                 ("cproj_w", "mlp.c_proj.weight"), ("in_proj_b", "attn.in_proj_bias"), ("out_proj_b", "attn.out_proj.bias"),
                 ("fc_b", "mlp.c_fc.bias"), ("cproj_b", "mlp.c_proj.bias"), ("ln1_w", "ln_1.weight"), ("ln1_b", "ln_1.bias"),
                 ("ln2_w", "ln_2.weight"), ("ln2_b", "ln_2.bias"))


def _block_struct(cls, tensors, i):
    """ctypes struct of block i's parameter (or gradient) pointers from a name -> tensor mapping / function."""
    p = f"transformer.resblocks.{i}."
    get = tensors if callable(tensors) else tensors.__getitem__
    return cls(**{field: get(p + name).data_ptr() for field, name in _BLOCK_FIELDS})


class _VitTrainFn(torch.autograd.Function):
    """forward = clip.model.VisionTransformer.forward layer by layer (activations kept);
    backward = its hand-written adjoint.  Replaces torch autograd through the reference's unfrozen
    ``PretrainedSSL`` encoder (lossyless/architectures.py:351-376)."""

    @staticmethod
    def forward(ctx, vit, x, names, *params):
        views = tuple(x) if isinstance(x, (tuple, list)) else (x,)
        R, P, Wd = vit.input_resolution, vit.patch_size, vit.width
        for v in views:
            N.require_cuda(v, "x")
            if v.dim() != 4 or v.shape[1] != 3 or v.shape[2] != R or v.shape[3] != R:
                raise ValueError(f"expected [B,3,{R},{R}] images, got {tuple(v.shape)}")
        x = views[0]
        B = sum(int(v.shape[0]) for v in views)
        if B == 0:
            raise ValueError("the fine-tuning path needs a non-empty batch")
        dt = vit._train_dtype
        w = vit._train_weights(dt)
        g = R // P
        S = g * g + 1
        M = B * S
        # im2col is pure data movement: [B,3,g,P,g,P] -> [B*g*g, 3*P*P]; ll_patchify does it together with the cast to the
        # 16-bit type, one view at a time into its slice (no float32 concatenation of the views, no separate cast)
        a_patch = torch.empty((B * g * g, 3 * P * P), dtype=dt, device=x.device)
        row = 0
        for v in views:
            if v.shape[0] == 0:
                continue
            if v.dtype in (torch.float32, torch.float16, torch.bfloat16) and 3 * P * P % 8 == 0:
                vc = v.contiguous()
                N.call("ll_patchify", N.ptr(vc), N.dtype_code(vc), N.ptr(a_patch[row:]), vc.shape[0], R, P, N.dtype_code(a_patch), N.stream())
            else:
                n = v.shape[0]
                a_patch[row: row + n * g * g] = v.to(dt).reshape(n, 3, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(n * g * g, 3 * P * P)
            row += v.shape[0] * g * g
        patch = _gemm(a_patch, w["conv1.weight"])
        # x = cat([class_embedding, conv1(x)], 1) + positional_embedding in one pass (no concatenated copy)
        tok = torch.empty((M, Wd), dtype=dt, device=x.device)
        N.call("ll_embed_tokens", N.ptr(patch), N.ptr(w["class_embedding"].contiguous()), N.ptr(w["positional_embedding"].contiguous()),
               N.ptr(tok), B, S, Wd, N.dtype_code(tok), N.stream())
        xs = _ln(tok, w["ln_pre.weight"], w["ln_pre.bias"])
        saved = []
        if NATIVE_BLOCKS:
            blocks = w.get("_blocks")
            if blocks is None:
                blocks = w["_blocks"] = [_block_struct(N.VitBlockParams, w, i) for i in range(vit.layers)]
            act_elems = int(N.load().ll_vit_block_act_elems(B, S, Wd))
            x2_off = int(N.load().ll_vit_block_x2_offset_elems(B, S, Wd))
            arena = torch.empty((vit.layers, act_elems), dtype=dt, device=x.device)  # every block's activation record
            x_in, code = xs, N.dtype_code(xs)
            cls_last = CLS_LAST_BLOCK
            for i in range(vit.layers):
                saved.append(x_in)
                if cls_last and i == vit.layers - 1:
                    # the block's output is the [B, width] class-token tensor itself
                    N.call("ll_vit_block_forward_cls", ctypes.byref(blocks[i]), N.ptr(x_in), N.ptr(arena[i]), B, S, vit.heads, Wd, code, N.stream())
                    off = int(N.load().ll_vit_block_cls_x2_offset_elems(B, S, Wd))
                    x_in = arena[i, off: off + B * Wd].view(B, Wd)
                else:
                    N.call("ll_vit_block_forward", ctypes.byref(blocks[i]), N.ptr(x_in), N.ptr(arena[i]), B, S, vit.heads, Wd, code, N.stream())
                    x_in = arena[i, x2_off: x2_off + M * Wd].view(M, Wd)
            saved.append(arena)
            xs = x_in
        for i in range(vit.layers if not NATIVE_BLOCKS else 0):
            p = f"transformer.resblocks.{i}."
            h1 = _ln(xs, w[p + "ln_1.weight"], w[p + "ln_1.bias"])
            qkv = _gemm(h1, w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"])
            a = torch.empty((M, Wd), dtype=dt, device=x.device)
            N.call("ll_attention", N.ptr(qkv), N.ptr(a), B, S, vit.heads, N.dtype_code(qkv), N.stream())
            x1 = _gemm(a, w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"], residual=xs)
            h2 = _ln(x1, w[p + "ln_2.weight"], w[p + "ln_2.bias"])
            u = _gemm(h2, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"])
            f = torch.empty_like(u)
            N.call("ll_quickgelu_forward", N.ptr(u), N.ptr(f), u.numel(), N.dtype_code(u), N.stream())
            x2 = _gemm(f, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"], residual=x1)
            saved.append((xs, h1, qkv, a, x1, h2, u, f))
            xs = x2
        cls_rows = NATIVE_BLOCKS and CLS_LAST_BLOCK  # xs is already x[:, 0, :]
        xc = xs if cls_rows else xs.view(B, S, Wd)[:, 0].contiguous()
        hc = _ln(xc, w["ln_post.weight"], w["ln_post.bias"])
        z = _gemm(hc, w["proj_t"])
        ctx.vit, ctx.names, ctx.w = vit, names, w
        ctx.state = dict(a_patch=a_patch, tok=tok, saved=saved, xc=xc, hc=hc, B=B, S=S, M=M, native=NATIVE_BLOCKS, cls_rows=cls_rows)
        ctx.needs = [p.requires_grad for p in params]
        return z.float()

    @staticmethod
    def backward(ctx, g_z):
        vit, w, st = ctx.vit, ctx.w, ctx.state
        dt = vit._train_dtype
        B, S, M, Wd = st["B"], st["S"], st["M"], vit.width
        dev = g_z.device
        # flat mode: every gradient is written where it will be all-reduced and consumed by the optimiser kernel
        # (16-bit for the GEMM-produced ones, float32 for the vectors); otherwise collected and returned to autograd
        flat = vit.flat_state() if all(ctx.needs) else None
        G = {}
        if flat is not None:
            flat.begin_backward()
            dst = flat.grad_view
        else:
            dst = lambda name: None  # noqa: E731
        gz = g_z.to(dt).contiguous()
        # z = hc @ proj
        g_hc = _gemm(gz, w["proj"])                      # [B, 768] = gz [B,512] @ proj[768,512]^T
        G["proj"] = _wgrad(st["hc"], gz, dst("proj"))    # [768, 512]
        g_xc, G["ln_post.weight"], G["ln_post.bias"] = _ln_backward(st["xc"], g_hc, w["ln_post.weight"], out_w=dst("ln_post.weight"), out_b=dst("ln_post.bias"))
        if st["cls_rows"]:
            g_x = torch.empty((M, Wd), dtype=dt, device=dev)  # written by the last block's backward
        else:
            g_x = torch.zeros((B, S, Wd), dtype=dt, device=dev)
            g_x[:, 0] = g_xc
            g_x = g_x.view(M, Wd)
        if st["native"]:
            blocks, arena = w["_blocks"], st["saved"][vit.layers]
            scratch = torch.empty(int(N.load().ll_vit_block_backward_scratch_elems(B, S, Wd)), dtype=dt, device=dev)
            code = N.dtype_code(g_x)
            for k, i in enumerate(reversed(range(vit.layers))):
                p = f"transformer.resblocks.{i}."
                if flat is not None:
                    grads = _block_struct(N.VitBlockGrads, flat.grad_view, i)
                else:
                    for _, name in _BLOCK_FIELDS:
                        shape = w[p + name].shape
                        G[p + name] = (torch.empty(shape, dtype=dt, device=dev) if _is_matrix(p + name)
                                       else torch.zeros(shape, dtype=torch.float32, device=dev))
                    grads = _block_struct(N.VitBlockGrads, G, i)
                if st["cls_rows"] and i == vit.layers - 1:
                    N.call("ll_vit_block_backward_cls", ctypes.byref(blocks[i]), N.ptr(st["saved"][i]), N.ptr(arena[i]), N.ptr(g_xc), N.ptr(g_x),
                           N.ptr(scratch), ctypes.byref(grads), B, S, vit.heads, Wd, code, N.stream())
                else:
                    N.call("ll_vit_block_backward", ctypes.byref(blocks[i]), N.ptr(st["saved"][i]), N.ptr(arena[i]), N.ptr(g_x), N.ptr(scratch),
                           ctypes.byref(grads), B, S, vit.heads, Wd, code, N.stream())
                if flat is not None:
                    flat.reduce_mat_slice(k)  # this block's four weight gradients (and proj, with the last block) are final
            st["saved"] = None
        for k, i in enumerate(reversed(range(vit.layers if not st["native"] else 0))):
            p = f"transformer.resblocks.{i}."
            xs, h1, qkv, a, x1, h2, u, f = st["saved"][i]
            wt = vit._train_weights_t(dt, i)
            # x2 = x1 + f @ c_proj^T + b
            g_f = _gemm(g_x, wt["c_proj"])
            G[p + "mlp.c_proj.weight"], G[p + "mlp.c_proj.bias"] = _wgrad_bias(g_x, f, dst(p + "mlp.c_proj.weight"), dst(p + "mlp.c_proj.bias"))
            g_u = torch.empty_like(u)
            N.call("ll_quickgelu_backward", N.ptr(u), N.ptr(g_f), N.ptr(g_u), u.numel(), N.dtype_code(u), N.stream())
            del g_f
            g_h2 = _gemm(g_u, wt["c_fc"])
            G[p + "mlp.c_fc.weight"], G[p + "mlp.c_fc.bias"] = _wgrad_bias(g_u, h2, dst(p + "mlp.c_fc.weight"), dst(p + "mlp.c_fc.bias"))
            del g_u
            g_x1, G[p + "ln_2.weight"], G[p + "ln_2.bias"] = _ln_backward(x1, g_h2, w[p + "ln_2.weight"], g_res=g_x, out_w=dst(p + "ln_2.weight"), out_b=dst(p + "ln_2.bias"))
            # x1 = x + a @ out_proj^T + b
            g_a = _gemm(g_x1, wt["out_proj"])
            G[p + "attn.out_proj.weight"], G[p + "attn.out_proj.bias"] = _wgrad_bias(g_x1, a, dst(p + "attn.out_proj.weight"), dst(p + "attn.out_proj.bias"))
            g_qkv = torch.empty_like(qkv)
            N.call("ll_attention_backward", N.ptr(qkv), N.ptr(g_a), N.ptr(g_qkv), B, S, vit.heads, N.dtype_code(qkv), N.stream())
            g_h1 = _gemm(g_qkv, wt["in_proj"])
            G[p + "attn.in_proj_weight"], G[p + "attn.in_proj_bias"] = _wgrad_bias(g_qkv, h1, dst(p + "attn.in_proj_weight"), dst(p + "attn.in_proj_bias"))
            del g_qkv, g_a
            g_x, G[p + "ln_1.weight"], G[p + "ln_1.bias"] = _ln_backward(xs, g_h1, w[p + "ln_1.weight"], g_res=g_x1, out_w=dst(p + "ln_1.weight"), out_b=dst(p + "ln_1.bias"))
            st["saved"][i] = None
            if flat is not None:
                flat.reduce_mat_slice(k)  # this block's four weight gradients (and proj, with the last block) are final
        g_tok, G["ln_pre.weight"], G["ln_pre.bias"] = _ln_backward(st["tok"], g_x, w["ln_pre.weight"], out_w=dst("ln_pre.weight"), out_b=dst("ln_pre.bias"))
        # adjoint of the token assembly, g_tok read once: the contiguous patch rows for conv1's weight-gradient GEMM and the
        # batch sums that are the embeddings' gradients (flat mode: accumulated into the zeroed float32 vector buffer)
        g_pos, g_cls = dst("positional_embedding"), dst("class_embedding")
        if g_pos is None:
            g_pos = G["positional_embedding"] = torch.zeros((S, Wd), dtype=torch.float32, device=dev)
            g_cls = G["class_embedding"] = torch.zeros(Wd, dtype=torch.float32, device=dev)
        g_patch = torch.empty((B * (S - 1), Wd), dtype=dt, device=dev)
        N.call("ll_embed_backward", N.ptr(g_tok), N.ptr(g_patch), N.ptr(g_pos), N.ptr(g_cls), B, S, Wd, N.dtype_code(g_tok), N.stream())
        conv_dst = dst("conv1.weight")
        G["conv1.weight"] = _wgrad(g_patch, st["a_patch"], conv_dst.view(Wd, -1) if conv_dst is not None else None).view(vit.conv1.weight.shape)
        if vit._native is not None:
            vit._native["stale"] = True  # an optimiser step is about to change the parameters
        ctx.state = None
        if flat is not None:
            flat.reduce_mat_slice(vit.layers)  # conv1
            flat.reduce_vec()
            return (None, None, None) + (None,) * len(ctx.names)
        grads = tuple(G[n].float() if need else None for n, need in zip(ctx.names, ctx.needs))
        return (None, None, None) + grads


def _train_weights(self, dt):
    """16-bit compute copies of the float32 master parameters (the autocast casts), rebuilt when an
    optimiser step bumped a parameter's version.  LayerNorm gains / biases and linear biases stay fp32.
    Flat mode: the matrices' copies are views of one flat 16-bit buffer that the optimiser kernel itself keeps up to
    date (``FlatTrainState.adamw_step``), the float32 vectors are the master views themselves -- no per-tensor casts."""
    flat = self.flat_state()
    versions = tuple(p._version for p in self.parameters())
    key = (dt, versions, self.proj.device, flat.step_count if flat is not None else -1)
    if self._train_cache is not None and self._train_cache["key"] == key:
        return self._train_cache["w"]
    w = {}
    with torch.no_grad():
        if flat is not None:
            flat.refresh_w16(versions)
            for n in flat.names:
                if _is_matrix(n):
                    w[n] = flat.view(flat.w16, n)
                elif n in ("class_embedding", "positional_embedding"):
                    w[n] = flat.view(flat.master, n).to(dt)
                else:
                    w[n] = flat.view(flat.master, n)
        else:
            for n, p in self.named_parameters():
                if n.endswith("bias") or "ln_" in n:
                    w[n] = p.detach().float().contiguous()
                else:
                    w[n] = p.detach().to(dt).contiguous()
        w["conv1.weight"] = w["conv1.weight"].reshape(self.width, -1)
        w["proj_t"] = w["proj"].t().contiguous()
    self._train_cache = dict(key=key, w=w, wt={})
    return w


def _train_weights_t(self, dt, i):
    """Transposed 16-bit copies of block i's four matrices: the W operands of the data-gradient GEMMs
    (dX = dY @ W is D = A W'^T with W' = W^T stored row-major)."""
    cache = self._train_cache["wt"]
    if i not in cache:
        w = self._train_cache["w"]
        p = f"transformer.resblocks.{i}."
        tr = lambda t: _transpose_pad(t)[:, : t.shape[0]].contiguous() if t.shape[0] % 64 else _transpose_pad(t)
        cache[i] = dict(in_proj=tr(w[p + "attn.in_proj_weight"]), out_proj=tr(w[p + "attn.out_proj.weight"]),
                        c_fc=tr(w[p + "mlp.c_fc.weight"]), c_proj=tr(w[p + "mlp.c_proj.weight"]))
    return cache[i]


VisionTransformer._train_weights = _train_weights
VisionTransformer._train_weights_t = _train_weights_t


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
