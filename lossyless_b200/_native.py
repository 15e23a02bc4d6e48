"""ctypes binding of liblossyless_b200.so (the C ABI in include/lossyless_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this
module raises.  Build it with ``python -m lossyless_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

import torch

from . import build as _build

LL_F32, LL_F16, LL_BF16 = 0, 1, 2
LL_EB_NPARAM = 58

_DTYPES = {torch.float32: LL_F32, torch.float16: LL_F16, torch.bfloat16: LL_BF16}


class NativeError(RuntimeError):
    pass


class VitConfig(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("width", "layers", "heads", "patch", "res", "out_dim", "dtype")]


class VitWeights(ctypes.Structure):
    _fields_ = [
        ("conv1_w", c_void_p), ("class_emb", c_void_p), ("pos_emb", c_void_p),
        ("ln_pre_w", c_void_p), ("ln_pre_b", c_void_p), ("ln_post_w", c_void_p), ("ln_post_b", c_void_p),
        ("proj_t", c_void_p),
        ("in_proj_w", POINTER(c_void_p)), ("in_proj_b", POINTER(c_void_p)),
        ("out_proj_w", POINTER(c_void_p)), ("out_proj_b", POINTER(c_void_p)),
        ("ln1_w", POINTER(c_void_p)), ("ln1_b", POINTER(c_void_p)),
        ("ln2_w", POINTER(c_void_p)), ("ln2_b", POINTER(c_void_p)),
        ("fc_w", POINTER(c_void_p)), ("fc_b", POINTER(c_void_p)),
        ("cproj_w", POINTER(c_void_p)), ("cproj_b", POINTER(c_void_p)),
    ]


class VitBlockParams(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("in_proj_w", "out_proj_w", "fc_w", "cproj_w", "in_proj_b", "out_proj_b", "fc_b", "cproj_b",
                                        "ln1_w", "ln1_b", "ln2_w", "ln2_b")]


class VitBlockGrads(ctypes.Structure):
    _fields_ = VitBlockParams._fields_


_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "ll_version": (c_int, []),
    "ll_last_error": (c_char_p, []),
    "ll_check_device": (c_int, [c_int]),
    "ll_pmf_to_quantized_cdf": (c_int, [_P, c_int, c_int, _P]),
    "ll_build_cdf_tables": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "ll_container_index": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P]),
    "ll_rans_stream_cap_bytes": (c_int, [c_int]),
    "ll_rans_set_impl": (c_int, [c_int]),
    "ll_rans_build_enc_table": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "ll_eb_rans_encode": (c_int, [_P, c_int, c_int64, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "ll_rans_encode_indexed": (c_int, [_P, c_int, c_int64, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "ll_rans_decode_indexed": (c_int, [_P, _P, _P, c_int64, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ll_rans_pack": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "ll_rans_decode": (c_int, [_P, _P, _P, c_int64, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ll_eb_pack_params": (c_int, [POINTER(_P), POINTER(_P), POINTER(_P), c_int, _P, _P]),
    "ll_eb_forward": (c_int, [_P, c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_float, _P, _P, _P, _P]),
    "ll_eb_backward": (c_int, [_P, c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ll_gc_forward": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P]),
    "ll_gc_backward": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P, _P, _P, _P]),
    "ll_vit_block_act_elems": (c_int64, [c_int64, c_int, c_int]),
    "ll_vit_block_x2_offset_elems": (c_int64, [c_int64, c_int, c_int]),
    "ll_vit_block_backward_scratch_elems": (c_int64, [c_int64, c_int, c_int]),
    "ll_vit_block_forward": (c_int, [POINTER(VitBlockParams), _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_vit_block_backward": (c_int, [POINTER(VitBlockParams), _P, _P, _P, _P, POINTER(VitBlockGrads), c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_vit_block_cls_x2_offset_elems": (c_int64, [c_int64, c_int, c_int]),
    "ll_vit_block_forward_cls": (c_int, [POINTER(VitBlockParams), _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_vit_block_backward_cls": (c_int, [POINTER(VitBlockParams), _P, _P, _P, _P, _P, POINTER(VitBlockGrads), c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_gather_cls_rows": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_scatter_cls_rows": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_sumsq_scratch_floats": (c_int, []),
    "ll_sumsq": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P]),
    "ll_adamw_flat": (c_int, [_P, _P, c_int, _P, _P, _P, c_int, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64, _P, _P]),
    "ll_cast_flat": (c_int, [_P, _P, c_int, c_int64, _P]),
    "ll_eb_unpack_grads": (c_int, [_P, POINTER(_P), POINTER(_P), c_int, POINTER(_P), POINTER(_P), POINTER(_P), _P]),
    "ll_eb_aux_loss": (c_int, [_P, _P, _P, c_int, _P, _P, _P]),
    "ll_vit_create": (c_int, [POINTER(VitConfig), POINTER(VitWeights), POINTER(_P)]),
    "ll_vit_destroy": (None, [_P]),
    "ll_vit_workspace_bytes": (c_int64, [_P, c_int64]),
    "ll_vit_forward": (c_int, [_P, _P, c_int, c_int64, _P, _P, c_int64, _P]),
    "ll_vit_launches": (c_int64, [_P, c_int64]),
    "ll_vit_set_chunk": (c_int, [_P, c_int64]),
    "ll_vit_set_fuse_layernorm": (c_int, [_P, c_int]),
    "ll_vit_profile_begin": (c_int, [_P]),
    "ll_vit_profile_end": (c_int, [_P, POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int64)]),
    "ll_vit_profile_breakdown": (c_int, [_P, POINTER(ctypes.c_double)]),
    "ll_layernorm": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, _P]),
    "ll_attention": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_attention_set_impl": (c_int, [c_int]),
    "ll_transpose_pad": (c_int, [_P, _P, c_int64, c_int, c_int64, _P]),
    "ll_transpose_pad_colsum": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, _P, _P]),
    "ll_colsum": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "ll_embed_tokens": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_embed_backward": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_quickgelu_forward": (c_int, [_P, _P, c_int64, c_int, _P]),
    "ll_quickgelu_backward": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "ll_layernorm_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "ll_attention_backward": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_attention_backward_set_impl": (c_int, [c_int]),
    "ll_gemm": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_clip_preprocess_set_impl": (c_int, [c_int]),
    "ll_gemm_tn": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_gemm_nn": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_gemm_dual": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_gemm_tn_workspace_bytes": (c_int64, [c_int, c_int]),
    "ll_gemm_tn_splitk": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P, c_int64, _P]),
    "ll_gemm_set_splitk": (c_int, [c_int]),
    "ll_vit_block_set_fusion": (c_int, [c_int]),
    "ll_gemm_patch_embed": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "ll_vit_set_direct_patch": (c_int, [_P, c_int]),
    "ll_vit_set_cls_last_block": (c_int, [_P, c_int]),
    "ll_patchify": (c_int, [_P, c_int, _P, c_int64, c_int, c_int, c_int, _P]),
    "ll_gemm_ln": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, _P]),
    "ll_gemm_set_tile_n": (c_int, [c_int]),
    "ll_gemm_set_cta_pair": (c_int, [c_int]),
    "ll_clip_preprocess": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, POINTER(c_float), POINTER(c_float),
                                  c_int, _P, _P, c_int, _P]),
    "ll_contrastive_logits": (c_int, [_P, c_int, c_int, _P, _P, _P, c_int, _P]),
    "ll_contrastive_logits_backward": (c_int, [_P, c_int, _P, _P, c_int, c_int, _P, _P]),
    "ll_contrastive_ce_forward": (c_int, [_P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P]),
    "ll_contrastive_ce_backward": (c_int, [_P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, c_int, _P, _P]),
    "ll_contrastive_ce_forward_dev": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "ll_contrastive_ce_backward_dev": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, c_int, _P, _P]),
    "ll_lossy_z_forward": (c_int, [_P, _P, c_int64, c_int, c_float, _P, _P]),
    "ll_lossy_z_backward": (c_int, [_P, _P, c_int64, c_int, c_float, _P, _P, _P]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load the shared library (never builds implicitly on a GPU box; fails loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(
            f"{path} is missing: the CUDA extension has not been built. Run "
            "`python -m lossyless_b200.build` (needs nvcc). There is no CPU fallback.")
    L = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().ll_last_error()
        raise NativeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero code."""
    check(getattr(load(), name)(*args), name)


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise NativeError(f"unsupported dtype {t.dtype}") from None


def ptr(t):
    """Device/host pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise NativeError("tensor handed to the native library must be contiguous")
    return c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """cudaStream_t of torch's current stream on the current device.  The raw C accessors are used when this
    torch build has them: ``torch.cuda.current_stream()`` costs ~13 us of Python per call, which was a quarter
    of the host time of a training step that makes ~300 native calls."""
    if _raw_stream is not None and _raw_device is not None:
        return c_void_p(_raw_stream(_raw_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise NativeError(
            f"{name} lives on {t.device}; the lossyless_b200 kernels only run on a B200 (cuda). "
            "There is no CPU fallback.")


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr
