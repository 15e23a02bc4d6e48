"""CPU: host-side mirror of the reference interface -- hub entry points, state-dict contract,
container helpers, device checks (no kernels are launched)."""
import io
import os
import struct

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hubconf_surface(native_lib):
    import hubconf

    assert {"clip_compressor_b001", "clip_compressor_b005", "clip_compressor_b01"} <= set(dir(hubconf))
    assert isinstance(hubconf.dependencies, list)
    comp, transform = hubconf.clip_compressor_b005(device="cpu")
    assert comp.z_dim == 512 and comp.side_z_dim == 102 and comp.device == "cpu"
    assert comp.scaling.shape == (512,) and comp.biasing.shape == (512,)
    assert comp.entropy_bottleneck._quantized_cdf.shape == (512, 32)
    assert not comp.training and callable(transform)
    for m in ("compress", "decompress", "get_rate", "compress_dataset", "decompress_dataset", "process_z_in",
              "process_z_out", "make_pickable_", "undo_pickable_", "forward"):
        assert callable(getattr(comp, m))
    # exp(scaling) ranges quoted in SURVEY 8a-2 for b005
    es = comp.scaling.exp()
    assert 4.9 < es.min() < es.max() < 6.5
    # reference behaviour: compressing on CPU is refused
    with pytest.raises(ValueError, match="only implemented on GPU"):
        comp.compress_dataset([], "/tmp/none.bin")
    # CLIP key names: a visual-tower checkpoint loads
    keys = set(comp.clip.state_dict())
    assert {"conv1.weight", "class_embedding", "positional_embedding", "proj", "ln_pre.weight", "ln_post.bias",
            "transformer.resblocks.11.attn.in_proj_weight", "transformer.resblocks.0.attn.out_proj.bias",
            "transformer.resblocks.3.mlp.c_fc.weight", "transformer.resblocks.3.mlp.c_proj.bias",
            "transformer.resblocks.5.ln_1.weight", "transformer.resblocks.5.ln_2.bias"} <= keys
    assert sum(v.numel() for v in comp.clip.state_dict().values()) == 87849216
    # process_z_in / out are inverse affine maps with the reference's shapes
    z = torch.randn(3, 512)
    assert comp.process_z_in(z).shape == (3, 512, 1, 1)
    assert torch.allclose(comp.process_z_out(comp.process_z_in(z)), z, atol=1e-5)


def test_preprocess_transform_matches_clip_recipe():
    from PIL import Image

    from lossyless_b200.clip_vit import CLIP_MEAN, CLIP_STD, clip_preprocess

    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 256, (96, 96, 3), dtype=np.uint8))  # STL10-shaped
    x = clip_preprocess(224)(img)
    assert x.shape == (3, 224, 224) and x.dtype == torch.float32
    lo = (0 - np.array(CLIP_MEAN)) / np.array(CLIP_STD)
    hi = (1 - np.array(CLIP_MEAN)) / np.array(CLIP_STD)
    for c in range(3):
        assert lo[c] - 1e-4 <= x[c].min() and x[c].max() <= hi[c] + 1e-4


def test_container_helpers_roundtrip():
    from lossyless_b200.hub.compressor import read_bytes, read_uints, write_bytes, write_uints

    f = io.BytesIO()
    write_uints(f, (3,))
    for b in (b"abcd", b"", b"12345678"):
        write_uints(f, (len(b),))
        write_bytes(f, b)
    assert f.getvalue()[:4] == struct.pack(">I", 3)
    f.seek(0)
    n = read_uints(f, 1)[0]
    got = [read_bytes(f, read_uints(f, 1)[0]) if True else None for _ in range(n)]
    assert got == [b"abcd", b"", b"12345678"]


def test_vit_requires_cuda(native_lib):
    from lossyless_b200._native import NativeError
    from lossyless_b200.clip_vit import VisionTransformer

    vit = VisionTransformer(224, 32, 768, 2, 12, 512)
    with pytest.raises(NativeError, match="no CPU fallback"):
        vit(torch.zeros(1, 3, 224, 224))


def test_downstream_linear_eval_runs():
    """tools/z_linear_eval.py (README.md:66-74 / utils/Z_linear_eval.py recipe) on separable synthetic Z."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("z_linear_eval", os.path.join(ROOT, "tools", "z_linear_eval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    centers = rng.normal(size=(5, 512)) * 2
    Y = rng.integers(0, 5, 600)
    Z = (centers[Y] + rng.normal(size=(600, 512))).astype(np.float32)
    res = mod.linear_eval(Z[:400], Y[:400], Z[400:], Y[400:])
    assert res["test/pred/acc"] > 0.95 and abs(res["test/pred/err"] - (1 - res["test/pred/acc"])) < 1e-12
