"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 paths -- sharding + container
merge (byte-identical to the single-rank file), the reference's gather-of-logit-tiles semantics
(forward order, local-only gradient) and the gradient bucket all-reduce."""
import os
import struct
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records(lo, hi):
    """fake per-image rANS records: deterministic bytes of varying (multiple of 4) length"""
    out = bytearray()
    for i in range(lo, hi):
        payload = bytes([(i * 7 + k) % 251 for k in range(8 + 4 * (i % 5))])
        out += struct.pack(">I", len(payload)) + payload
    return bytes(out)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lossyless_b200.helpers import gather_from_gpus
    from lossyless_b200.learnable_compressors import allreduce_mean_
    from lossyless_b200.parallel import gather_record_blobs, shard_range, write_container
    from oracle import losses as olosses

    # 1. sharded encode: contiguous slices, merged file == single-rank file
    n = 23
    lo, hi = shard_range(n, rank, world)
    blobs, counts = gather_record_blobs(_records(lo, hi), hi - lo)
    if rank == 0:
        write_container(os.path.join(tmp, "merged.bin"), blobs, counts)

    # 2. gather of logit tiles: own tile first, remote tiles constant (no gradient), as the oracle
    g = torch.Generator().manual_seed(100 + rank)
    tile = torch.randn(6, 6, generator=g, requires_grad=True)
    tiles = gather_from_gpus(tile)
    logits = torch.cat([tiles[rank]] + [t for r, t in enumerate(tiles) if r != rank], dim=1)
    other = torch.randn(6, 6, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    expect = olosses.contrastive_gather([tile.detach(), other] if rank == 0 else [other, tile.detach()], rank)
    assert torch.equal(logits.detach(), expect)
    w = torch.arange(72.0).reshape(6, 12)
    (logits * w).sum().backward()
    assert torch.equal(tile.grad, w[:, :6])  # only the local block's gradient comes back

    # 3. gradient bucket all-reduce (mean over ranks)
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(0.5), lin.bias.zero_()
    lin(torch.full((4, 3), float(rank + 1))).sum().backward()
    allreduce_mean_(list(lin.parameters()))
    # several buckets (bucket_bytes smaller than the gradients) and a 16-bit wire dtype give the same mean
    big = [torch.nn.Parameter(torch.zeros(n)) for n in (300, 7, 1200, 64)]
    for i, p in enumerate(big):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    allreduce_mean_(big, bucket_bytes=2048)
    for i, p in enumerate(big):
        assert torch.equal(p.grad, torch.full_like(p, 1.5 * (i + 1)))
    for i, p in enumerate(big):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    allreduce_mean_(big, comm_dtype=torch.float16, bucket_bytes=1024)
    for i, p in enumerate(big):
        assert p.grad.dtype == torch.float32 and torch.equal(p.grad, torch.full_like(p, 1.5 * (i + 1)))
    assert torch.allclose(lin.weight.grad, torch.full((2, 3), 4 * 1.5))
    assert torch.allclose(lin.bias.grad, torch.full((2,), 4.0))

    # 4. sharded decode: the REFERENCE-written container (tests/golden/ref_hub.npz) is indexed on every rank, split by
    # the prefix sum of the record lengths, each rank decodes only its range (host decoder injected: this is the
    # CPU test of the sharding logic), rows gathered on rank 0 in record order
    from lossyless_b200.hub.compressor import ClipCompressor
    from lossyless_b200.parallel import decompress_dataset_sharded, split_records_by_bytes
    from oracle import eb as oeb

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_hub.npz"))
    path = os.path.join(tmp, f"ref_{rank}.bin")
    with open(path, "wb") as f:
        f.write(g["container"].tobytes())
    sd = torch.load(os.path.join(ROOT, "lossyless_b200", "hub", "weights", "beta5e-02_factorized_rate.pt"), map_location="cpu", weights_only=True)
    ebo = oeb.split_state_dict(sd)
    tab = oeb.update(ebo)
    seen = []

    def host_decode(data, offsets, lens, lo, hi):
        seen.append((lo, hi))
        raw = data.numpy()
        strings = [raw[offsets[i]: offsets[i] + lens[i]].tobytes() for i in range(lo, hi)]
        if not strings:
            return torch.zeros((0, 512))
        zd, _ = oeb.decompress(ebo, tab, strings)
        return oeb.process_z_out(zd, sd["biasing"], sd["scaling"])

    class _Host:  # the container reader is host logic of the compressor; the decode itself is injected above
        z_dim, device = 512, "cpu"
        load_container = ClipCompressor.load_container

    Z = decompress_dataset_sharded(_Host(), path, decode_fn=host_decode)
    spans = split_records_by_bytes(np.diff(np.concatenate([[0], np.cumsum(g["lens5"])])), world)  # smoke of the helper
    assert len(spans) == world
    assert len(seen) == 1 and seen[0][1] - seen[0][0] < 20  # this rank decoded a strict part of the file
    if rank == 0:
        assert Z.shape == (20, 512) and np.array_equal(Z, g["Z_hat"])
    else:
        assert Z is None
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
        merged = open(os.path.join(tmp, "merged.bin"), "rb").read()
    assert merged == struct.pack(">I", 23) + _records(0, 23)


def test_split_records_by_bytes():
    from lossyless_b200.parallel import split_records_by_bytes

    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 64, 5000):
        lens = 8 + 4 * rng.integers(0, 120, n)
        for world in (1, 2, 3, 8):
            spans = split_records_by_bytes(lens, world)
            assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(lo <= hi for lo, hi in spans)
            if n >= 64:  # balanced by bytes: no rank more than one record's worth away from the mean
                per = [int((lens[lo:hi] + 4).sum()) for lo, hi in spans]
                assert max(per) - min(per) <= 2 * (int(lens.max()) + 4)
    # skewed lengths: one huge record does not drag its neighbours along
    assert split_records_by_bytes([8, 4000, 8, 8], 2) == [(0, 1), (1, 4)]


def test_shard_range_covers_everything():
    from lossyless_b200.parallel import shard_range

    for n in (0, 1, 7, 8, 23, 5000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
