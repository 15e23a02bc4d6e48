"""GPU: the reference-facing hub API end to end (ViT -> bottleneck -> rANS -> container file)."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import eb as oeb

pytestmark = pytest.mark.gpu


class _Synthetic(torch.utils.data.Dataset):
    """STL10-shaped uint8 images pushed through the CLIP transform (BASELINE config 2)."""

    def __init__(self, n, transform):
        from PIL import Image

        rng = np.random.default_rng(0)
        self.items = [(transform(Image.fromarray(rng.integers(0, 256, (96, 96, 3), dtype=np.uint8))), i % 10) for i in range(n)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


@pytest.fixture(scope="module")
def compressor(cuda_device):
    import hubconf

    return hubconf.clip_compressor_b005(device="cuda")


def test_compress_decompress_roundtrip(compressor, weights_dir):
    comp, _ = compressor
    x = torch.randn(37, 3, 224, 224, generator=torch.Generator().manual_seed(0)).half().cuda()
    strings = comp.compress(x)
    assert len(strings) == 37 and all(isinstance(s, bytes) and len(s) % 4 == 0 for s in strings)
    z_hat = comp.decompress(strings)
    assert z_hat.shape == (37, 512) and z_hat.dtype == torch.float32
    assert torch.equal(z_hat, comp(x))  # forward(X) == decompress(compress(X)), as in the reference
    assert comp.get_rate(x) == pytest.approx(8 * np.mean([len(s) for s in strings]))
    # same z through the CPU oracle gives the same bytes (EB + rANS stage is bit exact given z)
    z = comp.clip(x)
    sd = torch.load(os.path.join(weights_dir, "beta5e-02_factorized_rate.pt"), map_location="cpu", weights_only=True)
    ebo = oeb.split_state_dict(sd)
    ref, _ = oeb.compress(ebo, oeb.update(ebo), oeb.process_z_in(z.cpu(), sd["biasing"], sd["scaling"]))
    assert strings == ref
    assert abs((z_hat.cpu() - z.float().cpu())).max() < 0.5 / sd["scaling"].exp().min() + 1e-3


def test_compress_dataset_file_format(compressor, tmp_path, capsys):
    comp, transform = compressor
    ds = _Synthetic(70, transform)
    f, lf = tmp_path / "Z.bin", tmp_path / "Y.npy"
    comp.compress_dataset(ds, f, label_file=lf, kwargs_dataloader=dict(batch_size=32, num_workers=0))
    out = capsys.readouterr().out
    assert "Rate:" in out and "bits/img | Encoding:" in out and "img/sec" in out
    blob = f.read_bytes()
    assert struct.unpack(">I", blob[:4])[0] == 70
    # byte-identical to the reference's writer fed with compress() outputs (hub/compressor.py:192-196)
    X = torch.stack([ds[i][0] for i in range(70)]).half().cuda()
    expect = struct.pack(">I", 70) + b"".join(struct.pack(">I", len(s)) + s for s in comp.compress(X))
    assert blob == expect
    Z, Y = comp.decompress_dataset(f, label_file=lf)
    assert "Decoding:" in capsys.readouterr().out
    assert Z.shape == (70, 512) and Z.dtype == np.float32 and Y.dtype == np.int64
    assert Y.tolist() == [i % 10 for i in range(70)]
    assert np.array_equal(Z, comp(X).cpu().numpy())
    # corrupt container: truncated file raises
    (tmp_path / "bad.bin").write_bytes(blob[:-5])
    with pytest.raises(ValueError):
        comp.decompress_dataset(tmp_path / "bad.bin")


def test_compress_stream_matches_blocking_calls(compressor):
    """compress_dataset's pipelined loop (batch i+1 enqueued before batch i is read back, staging slots shared
    across batches) returns, batch by batch, the bytes of the one-call-per-batch path -- ragged batch sizes,
    fp16 / fp32 / raw uint8 inputs."""
    comp, _ = compressor
    g = torch.Generator().manual_seed(11)
    batches = [torch.randn(n, 3, 224, 224, generator=g).half().pin_memory() for n in (5, 400, 1, 379)]
    batches.append(torch.randn(7, 3, 224, 224, generator=g).pin_memory())
    want = [bytes(comp.compress_batch_to_host(b)[0].numpy().data) for b in batches]
    got = [(bytes(rec.numpy().data), n) for rec, n in comp.compress_stream(iter(batches), sub_batch=128)]
    assert [n for _, n in got] == [b.shape[0] for b in batches]
    assert [r for r, _ in got] == want
    raw = [torch.randint(0, 256, (n, 96, 96, 3), dtype=torch.uint8, generator=g).pin_memory() for n in (3, 130)]
    want = [bytes(comp.compress_batch_to_host(b)[0].numpy().data) for b in raw]
    assert [bytes(rec.numpy().data) for rec, _ in comp.compress_stream(iter(raw))] == want
    assert list(comp.compress_stream(iter([]))) == []
    # coalescing: consecutive small batches share one encoder pass; one yield per group, records in input order
    small = [torch.randn(n, 3, 224, 224, generator=g).half().pin_memory() for n in (100, 100, 100, 60, 300, 7)]
    per_batch = [bytes(comp.compress_batch_to_host(b)[0].numpy().data) for b in small]
    got = [(bytes(rec.numpy().data), n) for rec, n in comp.compress_stream(iter(small), coalesce=256)]
    assert [n for _, n in got] == [200, 160, 300, 7]  # 100+100 | 100+60 | 300 (>= limit: alone) | 7
    assert b"".join(r for r, _ in got) == b"".join(per_batch)
    # default sub_batch (one H2D copy and one encoder pass per batch): the staging slots grow with the batch size
    # (5 -> 400) and are re-created while the previous batch may still be in flight
    got = [bytes(rec.numpy().data) for rec, _ in comp.compress_stream(iter(batches[:4]))]
    assert got == [bytes(comp.compress_batch_to_host(b)[0].numpy().data) for b in batches[:4]]
