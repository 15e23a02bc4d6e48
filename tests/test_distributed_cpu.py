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


def test_shard_range_covers_everything():
    from lossyless_b200.parallel import shard_range

    for n in (0, 1, 7, 8, 23, 5000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
