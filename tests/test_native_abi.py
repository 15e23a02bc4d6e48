"""CPU: the C-ABI library loads, exports every symbol include/lossyless_b200.h declares, and its
host-side (load-time) functions agree bit for bit with the oracle.  No device calls here."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch

from oracle import eb as oeb
from oracle import rans

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]


def test_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "lossyless_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ll_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 29
    from lossyless_b200 import _native

    for name in sorted(declared):
        assert hasattr(native_lib, name), f"{name} declared in the header but not exported"
        assert name in _native._SIGNATURES, f"{name} has no ctypes signature"
    assert set(_native._SIGNATURES) <= declared
    assert native_lib.ll_version() >= 100


def test_no_torch_types_and_sm100a_only(native_lib):
    from lossyless_b200 import build

    flags = " ".join(build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "sm_90" not in flags
    deps = os.popen(f"ldd {build.LIB}").read()
    assert "torch" not in deps and "c10" not in deps


def test_pmf_to_quantized_cdf_matches_oracle(native_lib):
    rng = np.random.default_rng(0)
    for n in (2, 3, 17, 33, 64):
        for _ in range(20):
            p = rng.random(n).astype(np.float32) ** 4
            p[rng.random(n) < 0.3] = 0.0
            p[0] = max(p[0], 1e-3)
            p = (p / p.sum()).astype(np.float32)
            out = np.zeros(n + 1, dtype=np.int32)
            rc = native_lib.ll_pmf_to_quantized_cdf(p.ctypes.data, n, 16, out.ctypes.data)
            assert rc == 0, native_lib.ll_last_error()
            assert np.array_equal(out, rans.pmf_to_quantized_cdf(p))
    bad = np.array([0.5, float("nan")], np.float32)
    out = np.zeros(3, np.int32)
    assert native_lib.ll_pmf_to_quantized_cdf(bad.ctypes.data, 2, 16, out.ctypes.data) == -3
    assert b"not finite" in native_lib.ll_last_error()


@pytest.mark.parametrize("tag", TAGS)
def test_entropy_bottleneck_update_matches_golden(native_lib, weights_dir, golden_dir, tag):
    """Product host path (torch CPU pmf + ll_build_cdf_tables) == oracle golden tables."""
    from lossyless_b200.entropy_models import EntropyBottleneck

    sd = torch.load(os.path.join(weights_dir, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)
    sd = {k[len("entropy_bottleneck."):]: v for k, v in sd.items() if k.startswith("entropy_bottleneck.")}
    eb = EntropyBottleneck(512, init_scale=10, filters=[3, 3, 3, 3])
    eb.load_state_dict(sd)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        eb._check_cdf_size()
    assert eb.update() is True and eb.update() is False and eb.update(force=True) is True
    g = np.load(os.path.join(golden_dir, "cdf_tables.npz"))
    assert np.array_equal(eb._quantized_cdf.numpy(), g[f"{tag}/cdf"])
    assert np.array_equal(eb._cdf_length.numpy(), g[f"{tag}/cdf_length"])
    assert np.array_equal(eb._offset.numpy(), g[f"{tag}/offset"])
    # aux loss (plain torch, 3 points per channel) equals the oracle's
    assert abs(eb.loss().item() - oeb.aux_loss(oeb.split_state_dict({f"entropy_bottleneck.{k}": v for k, v in sd.items()})).item()) < 1e-2
    # a checkpoint that already carries tables round-trips through state_dict (dynamic buffer sizes)
    eb2 = EntropyBottleneck(512, init_scale=10, filters=[3, 3, 3, 3])
    eb2.load_state_dict(eb.state_dict())
    assert np.array_equal(eb2._quantized_cdf.numpy(), g[f"{tag}/cdf"]) and eb2.update() is False


def test_entropy_bottleneck_param_count_and_errors(native_lib):
    from lossyless_b200.entropy_models import EntropyBottleneck

    eb = EntropyBottleneck(512, init_scale=10, filters=[3, 3, 3, 3])
    assert sum(p.numel() for p in eb.parameters()) == 31232  # "31.2 K" in notebooks/minimal_code.ipynb
    with pytest.raises(ValueError):
        eb.quantize(torch.zeros(2), "bogus")
    with pytest.raises(ValueError):
        EntropyBottleneck(4, entropy_coder="range")
    # no CPU fallback: the likelihood and the codec refuse CPU tensors loudly
    from lossyless_b200._native import NativeError

    eb.update()
    with pytest.raises(NativeError, match="no CPU fallback"):
        eb(torch.zeros(2, 512, 1, 1))
    with pytest.raises(NativeError, match="no CPU fallback"):
        eb.compress(torch.zeros(2, 512, 1, 1))


def test_container_index(native_lib):
    recs = [b"\x01\x02\x03\x04" * k for k in (2, 5, 3)]
    blob = struct.pack(">I", len(recs)) + b"".join(struct.pack(">I", len(r)) + r for r in recs)
    buf = np.frombuffer(blob, dtype=np.uint8)
    off = np.zeros(4, np.int64)
    lens = np.zeros(3, np.int32)
    assert native_lib.ll_container_index(buf.ctypes.data, buf.size, 4, 3, off.ctypes.data, lens.ctypes.data) == 0
    assert lens.tolist() == [8, 20, 12] and off.tolist() == [8, 20, 44, 56]
    for i, r in enumerate(recs):
        assert blob[off[i]: off[i] + lens[i]] == r
    # truncation is an error, not a crash
    assert native_lib.ll_container_index(buf.ctypes.data, buf.size - 3, 4, 3, off.ctypes.data, lens.ctypes.data) == -4
    assert b"truncated" in native_lib.ll_last_error()


def test_bad_arguments_are_errors_not_crashes(native_lib):
    assert native_lib.ll_rans_stream_cap_bytes(512) == 4 * ((512 * 52 + 31) // 32 + 4)
    assert native_lib.ll_rans_stream_cap_bytes(0) == -1
    assert native_lib.ll_gemm(None, None, None, None, None, 1, 128, 64, 1, 0, None) == -1
    assert native_lib.ll_eb_rans_encode(None, 0, 1, 512, None, None, None, None, 32, None, None, None, None, None, None) == -1
    assert native_lib.ll_vit_workspace_bytes(None, 8) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from lossyless_b200 import _native, build

    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(build, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        _native.load()
