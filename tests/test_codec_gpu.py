"""GPU parity: fused quantise + rANS encode / pack / decode kernels vs the CPU oracle.
Bit-exact (integer / byte work)."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import eb as oeb
from oracle import rans

pytestmark = pytest.mark.gpu
TAGS = ["beta1e-01", "beta5e-02", "beta1e-02"]


def _module(weights_dir, tag, device):
    from lossyless_b200.entropy_models import EntropyBottleneck

    sd = torch.load(os.path.join(weights_dir, f"{tag}_factorized_rate.pt"), map_location="cpu", weights_only=True)
    eb = EntropyBottleneck(512, init_scale=10, filters=[3, 3, 3, 3])
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items() if k.startswith("entropy_bottleneck.")})
    eb.update()
    return eb.to(device), sd, sd["biasing"].to(device), sd["scaling"].exp().to(device)


@pytest.mark.parametrize("tag", TAGS)
def test_golden_bytes(cuda_device, weights_dir, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "rans_vectors.npz"))
    eb, sd, b, es = _module(weights_dir, tag, cuda_device)
    z = torch.from_numpy(g[f"{tag}/z_f16"]).to(cuda_device)
    strings = eb.compress(z, bias=b, exp_scale=es)
    lens = g[f"{tag}/lens"]
    off = np.concatenate([[0], np.cumsum(lens)])
    gold = g[f"{tag}/bytes"].tobytes()
    assert [len(s) for s in strings] == lens.tolist()
    for i, s in enumerate(strings):
        assert s == gold[off[i]: off[i + 1]], f"row {i} differs"
    # symbols + device-side container records
    packed, offsets, lens_d, sym = eb.compress_to_device(z, b, es, mode=1, want_symbols=True)
    assert np.array_equal(sym.cpu().numpy(), g[f"{tag}/symbols"])
    rec = packed[: int(offsets[-1])].cpu().numpy().tobytes()
    assert rec == b"".join(struct.pack(">I", len(s)) + s for s in strings)
    # decode: symbols and z_hat (float32 ops are IEEE-exact -> bit-equal to the oracle)
    z_hat = eb.decompress(strings, [1, 1], bias=b, exp_scale=es).reshape(len(strings), 512)
    assert np.array_equal(z_hat.cpu().numpy(), g[f"{tag}/z_hat"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_random_vs_oracle_all_dtypes(cuda_device, weights_dir, dtype):
    eb, sd, b, es = _module(weights_dir, "beta1e-02", cuda_device)
    gen = torch.Generator().manual_seed(3)
    z = (torch.randn(1000 + 37, 512, generator=gen) * 0.6).to(dtype)  # ragged vs the 32-row CTA
    z[5] *= 30
    strings = eb.compress(z.to(cuda_device), bias=b, exp_scale=es)
    ebo = oeb.split_state_dict(sd)
    tab = oeb.update(ebo)
    ref, sym = oeb.compress(ebo, tab, oeb.process_z_in(z, sd["biasing"], sd["scaling"]))
    assert strings == ref
    # z_in fed directly (bias/exp_scale omitted) codes the same symbols
    z_in = oeb.process_z_in(z, sd["biasing"], sd["scaling"])
    assert eb.compress(z_in.to(cuda_device)) == ref
    _, dsym = eb.decompress_from_device(*_to_dev(ref, cuda_device), want_symbols=True)
    assert np.array_equal(dsym.cpu().numpy(), sym)


def _to_dev(strings, device):
    lens = np.array([len(s) for s in strings], np.int32)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    data = torch.frombuffer(bytearray(b"".join(strings)) + bytearray(16), dtype=torch.uint8).to(device)
    return data, torch.from_numpy(off).to(device), torch.from_numpy(lens).to(device)


def test_edge_cases(cuda_device, weights_dir):
    eb, sd, b, es = _module(weights_dir, "beta5e-02", cuda_device)
    assert eb.compress(torch.zeros(0, 512, device=cuda_device), bias=b, exp_scale=es) == []
    one = eb.compress(torch.zeros(1, 512, device=cuda_device), bias=b, exp_scale=es)
    assert len(one) == 1 and len(one[0]) % 4 == 0 and len(one[0]) >= 8
    # fp16 extremes: every symbol escapes with a 6-nibble raw value; still inside the scratch slot
    big = torch.full((3, 512), 65504.0, dtype=torch.float16, device=cuda_device)
    big[1] = -65504.0
    s = eb.compress(big, bias=b, exp_scale=es)
    cap = rans.stream_cap_bytes(512)
    assert all(len(x) <= cap for x in s)
    ebo = oeb.split_state_dict(sd)
    ref, _ = oeb.compress(ebo, oeb.update(ebo), oeb.process_z_in(big.cpu(), sd["biasing"], sd["scaling"]))
    assert s == ref
    # truncated / misaligned streams raise, they do not crash
    with pytest.raises(ValueError):
        eb.decompress([one[0][:8]], [1, 1], bias=b, exp_scale=es)
    with pytest.raises(ValueError):
        eb.decompress([one[0][:-2]], [1, 1], bias=b, exp_scale=es)
    # a model whose tables were never built says so (compressai's message)
    from lossyless_b200.entropy_models import EntropyBottleneck

    fresh = EntropyBottleneck(512).to(cuda_device)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        fresh.compress(torch.zeros(1, 512, device=cuda_device))


def test_full_size_roundtrip_property(cuda_device, weights_dir):
    """BASELINE sweep size (65 536 rows x 512): decode(encode(z)) reproduces the quantised symbols,
    computed independently with torch elementwise ops; total bits ~ the expected entropy."""
    eb, sd, b, es = _module(weights_dir, "beta5e-02", cuda_device)
    gen = torch.Generator(device=cuda_device).manual_seed(11)
    z = (torch.randn(65536, 512, generator=gen, device=cuda_device) * 0.45).half()
    packed, offsets, lens, sym = eb.compress_to_device(z, b, es, mode=0, want_symbols=True)
    med = eb._get_medians().reshape(-1)
    expect = torch.round((z.float() + b) * es - med).int()
    assert torch.equal(sym, expect)
    z_hat, dsym = eb.decompress_from_device(packed, offsets, lens, bias=b, exp_scale=es, want_symbols=True)
    assert torch.equal(dsym, expect)
    assert torch.equal(z_hat, (expect.float() + med) / es - b)
    bits = 8 * lens.float().mean().item()
    assert 1950 < bits < 2150, bits  # ~2055 bits/img for b005 on N(0, 0.45^2)
    assert int(offsets[-1]) == int(lens.sum())


@pytest.fixture
def codec_impl():
    """Sets the codec kernel choice (include/lossyless_b200.h: ll_rans_set_impl) and restores 'auto'."""
    from lossyless_b200 import _native as N

    yield lambda impl: N.call("ll_rans_set_impl", impl)
    N.call("ll_rans_set_impl", 0)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_stream_per_lane_kernels_vs_oracle(cuda_device, weights_dir, codec_impl, tag, dtype):
    """The throughput kernels (one stream per lane) forced on at a ragged row count: bytes, symbols and
    decoded values identical to the oracle's, i.e. to the stream-per-warp kernels'."""
    eb, sd, b, es = _module(weights_dir, tag, cuda_device)
    gen = torch.Generator().manual_seed(5)
    z = (torch.randn(3 * 128 + 32 + 5, 512, generator=gen) * 0.6).to(dtype)  # full CTAs, a full warp, a ragged warp
    z[7] *= 30   # a row of escapes
    z[40, ::3] = -200.0
    codec_impl(2)
    packed, offsets, lens, sym = eb.compress_to_device(z.to(cuda_device), b, es, mode=1, want_symbols=True)
    strings = eb.compress(z.to(cuda_device), bias=b, exp_scale=es)
    ebo = oeb.split_state_dict(sd)
    tab = oeb.update(ebo)
    ref, rsym = oeb.compress(ebo, tab, oeb.process_z_in(z, sd["biasing"], sd["scaling"]))
    assert strings == ref
    assert np.array_equal(sym.cpu().numpy(), rsym)
    rec = packed[: int(offsets[-1])].cpu().numpy().tobytes()
    assert rec == b"".join(struct.pack(">I", len(s)) + s for s in ref)
    z_hat, dsym = eb.decompress_from_device(*_to_dev(ref, cuda_device), bias=b, exp_scale=es, want_symbols=True)
    assert np.array_equal(dsym.cpu().numpy(), rsym)
    zh_ref, _ = oeb.decompress(ebo, tab, ref)
    assert torch.equal(z_hat.cpu(), oeb.process_z_out(zh_ref, sd["biasing"], sd["scaling"]))
    # same through the stream-per-warp kernels
    codec_impl(1)
    assert eb.compress(z.to(cuda_device), bias=b, exp_scale=es) == ref


def test_stream_per_lane_edge_cases(cuda_device, weights_dir, codec_impl):
    eb, sd, b, es = _module(weights_dir, "beta5e-02", cuda_device)
    codec_impl(2)
    one = eb.compress(torch.zeros(1, 512, device=cuda_device), bias=b, exp_scale=es)
    codec_impl(1)
    assert one == eb.compress(torch.zeros(1, 512, device=cuda_device), bias=b, exp_scale=es)
    codec_impl(2)
    big = torch.full((3, 512), 65504.0, dtype=torch.float16, device=cuda_device)
    big[1] = -65504.0
    s = eb.compress(big, bias=b, exp_scale=es)
    ebo = oeb.split_state_dict(sd)
    ref, _ = oeb.compress(ebo, oeb.update(ebo), oeb.process_z_in(big.cpu(), sd["biasing"], sd["scaling"]))
    assert s == ref
    assert torch.equal(eb.decompress(s, [1, 1], bias=b, exp_scale=es).reshape(3, 512).cpu(),
                       oeb.process_z_out(oeb.decompress(ebo, oeb.update(ebo), ref)[0], sd["biasing"], sd["scaling"]))
    with pytest.raises(ValueError):
        eb.decompress([one[0][:8]], [1, 1], bias=b, exp_scale=es)
    with pytest.raises(ValueError):
        eb.decompress([one[0][:-2]], [1, 1], bias=b, exp_scale=es)
    # a channel count that is not a multiple of the 64-channel chunk (config 4: C = 128 is; 102 is not)
    from lossyless_b200.entropy_models import EntropyBottleneck

    torch.manual_seed(0)
    small = EntropyBottleneck(102, init_scale=10, filters=[3, 3, 3, 3])
    small.update(force=True)
    small = small.to(cuda_device)
    z = torch.randn(77, 102, device=cuda_device) * 4
    got = small.compress(z)
    codec_impl(1)
    assert got == small.compress(z)
    codec_impl(2)
    assert torch.equal(small.decompress(got, [1, 1]).reshape(77, 102), small.quantize(z, "dequantize", small._get_medians().reshape(1, -1)))


def test_lane_decoder_quotient_paths(cuda_device, weights_dir, codec_impl):
    """The stream-per-lane decoder (csrc/codec.cu: rans_decode_lanes2_kernel) divides by exp(scaling) through a
    per-channel refined reciprocal when every dividend of the channel provably stays inside the exponent window of
    that sequence, and by the IEEE division otherwise: integer / zero / almost-integer medians and extreme scales
    take the second path.  Symbols equal the encoder's, values are bit-equal to torch's float32 division, and the
    stream-per-warp and stream-per-lane decoders agree (as does the default choice)."""
    from lossyless_b200.entropy_models import EntropyBottleneck

    sd = torch.load(os.path.join(weights_dir, "beta5e-02_factorized_rate.pt"), map_location="cpu", weights_only=True)
    ebsd = {k[len("entropy_bottleneck."):]: v.clone() for k, v in sd.items() if k.startswith("entropy_bottleneck.")}
    for c, m in ((0, 0.0), (1, 2.0), (2, -1.0), (3, 1e-25), (4, -3e-20), (5, 0.5), (6, 1.0 - 2.0 ** -24)):
        ebsd["quantiles"][c, 0, :] += m - ebsd["quantiles"][c, 0, 1]
        ebsd["quantiles"][c, 0, 1] = m   # the value itself, not a rounded sum
    eb = EntropyBottleneck(512, init_scale=10, filters=[3, 3, 3, 3])
    eb.load_state_dict(ebsd)
    eb.update(force=True)
    eb = eb.to(cuda_device)
    b = sd["biasing"].to(cuda_device).clone()
    scaling = sd["scaling"].clone()
    scaling[7], scaling[8], scaling[9], scaling[10] = 30.0, -30.0, 27.0, -27.0   # exp(+-30) leaves [2^-40, 2^40], exp(+-27) does not
    es = scaling.exp().to(cuda_device)
    med = eb._get_medians().reshape(-1)
    gen = torch.Generator(device=cuda_device).manual_seed(21)
    z = (torch.randn(4096 + 37, 512, generator=gen, device=cuda_device) * 0.6).half()
    z[7] *= 30
    z[100:110, :12] = 0.0
    outs = []
    for impl in (1, 2, 0):
        codec_impl(impl)
        packed, offsets, lens, sym = eb.compress_to_device(z, b, es, mode=0, want_symbols=True)
        z_hat, dsym = eb.decompress_from_device(packed, offsets, lens, bias=b, exp_scale=es, want_symbols=True)
        assert torch.equal(dsym, sym), impl
        assert torch.equal(z_hat, (sym.float() + med) / es - b), impl
        outs.append((packed[: int(lens.sum())].clone(), z_hat))
        # without the affine map: float(symbol) + median
        z_in_hat, _ = eb.decompress_from_device(packed, offsets, lens)
        assert torch.equal(z_in_hat, sym.float() + med), impl
    assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])


def test_large_pack_matches_single_cta_pack(cuda_device, weights_dir, codec_impl):
    """> 4096 rows take the three-launch pack (block sums in the offsets array): identical records."""
    eb, sd, b, es = _module(weights_dir, "beta1e-02", cuda_device)
    gen = torch.Generator(device=cuda_device).manual_seed(2)
    z = (torch.randn(4096 + 1024 + 17, 512, generator=gen, device=cuda_device) * 0.45).half()
    codec_impl(2)
    packed, offsets, lens, _ = eb.compress_to_device(z, b, es, mode=1)
    codec_impl(1)
    total = 0
    for lo, hi in ((0, 4096), (4096, z.shape[0])):  # pieces small enough for the single-CTA pack
        p2, o2, l2, _ = eb.compress_to_device(z[lo:hi], b, es, mode=1)
        n = int(o2[-1])
        assert torch.equal(packed[total: total + n], p2[:n])
        assert torch.equal(offsets[lo:hi] - total, o2[:-1])
        assert torch.equal(lens[lo:hi], l2)
        total += n
    assert int(offsets[-1]) == total
