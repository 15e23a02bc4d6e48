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


def test_downstream_probe_on_decompressed_representations(cuda_device, tmp_path, weights_dir):
    """the consumer side.  Class-structured synthetic images -> compress_dataset -> decompress_dataset ->
    the reference's ``Predictor`` probe (lossyless/predictors.py) trained on Z_hat.  The same probe trained on the
    representations of the float32 CPU oracle (same weights, same bottleneck, no 16-bit encoder, no codec) reaches the
    same accuracy: the fp16 tower's symbol flips do not move the downstream task."""
    import hubconf
    from lossyless_b200.helpers import NamespaceMap
    from lossyless_b200.predictors import Predictor
    from oracle import vit as ovit

    sd_vit = ovit.init_state_dict(0)
    comp, _ = hubconf.clip_compressor_b005(device="cuda", clip_state_dict=sd_vit)
    g = torch.Generator().manual_seed(0)
    n_cls, per = 6, 40
    templates = torch.randn(n_cls, 3, 224, 224, generator=g)
    y = torch.arange(n_cls * per) % n_cls
    X = (templates[y] + 1.0 * torch.randn(n_cls * per, 3, 224, 224, generator=g)).half()
    ds = torch.utils.data.TensorDataset(X, y)
    comp.compress_dataset(ds, tmp_path / "Z.bin", label_file=tmp_path / "Y.npy", kwargs_dataloader=dict(batch_size=64, num_workers=0), is_info=False)
    Z_hat, Y = comp.decompress_dataset(tmp_path / "Z.bin", label_file=tmp_path / "Y.npy", is_info=False)
    assert Z_hat.shape == (n_cls * per, 512) and np.array_equal(Y, y.numpy())
    with torch.no_grad():  # the oracle's float32 representations, quantised by the oracle bottleneck
        sd = torch.load(os.path.join(weights_dir, "beta5e-02_factorized_rate.pt"), map_location="cpu", weights_only=True)
        ebo = oeb.split_state_dict(sd)
        z32 = torch.cat([ovit.forward({k: v.half().float() for k, v in sd_vit.items()}, X[i:i + 60].float()) for i in range(0, len(X), 60)])
        zh_in, _ = oeb.forward(ebo, oeb.process_z_in(z32, sd["biasing"], sd["scaling"]))
        Z_ref = oeb.process_z_out(zh_in, sd["biasing"], sd["scaling"])

    def probe_acc(Z):
        hp = NamespaceMap(stage="pred", data=dict(shape=(512,), target_shape=(n_cls,), target_is_clf=True, balancing_weights={}),
                          predictor=dict(arch="linear", arch_kwargs={}),
                          optimizer_pred=dict(mode="Adam", kwargs=dict(lr=3e-2)), scheduler_pred=None)
        torch.manual_seed(0)
        pred = Predictor(hp)
        Z = torch.as_tensor(Z)
        tr = torch.arange(len(Z)) % 4 != 0  # 3/4 train, 1/4 test
        pred.fit([(Z[tr], y[tr])], epochs=60)
        return pred.evaluate([(Z[~tr], y[~tr])])["acc"]

    acc_gpu, acc_ref = probe_acc(Z_hat), probe_acc(Z_ref.numpy())
    assert acc_ref > 0.9 and abs(acc_gpu - acc_ref) <= 0.05, (acc_gpu, acc_ref)
    assert float((torch.as_tensor(Z_hat) != Z_ref).float().mean()) < 0.05  # a few percent of the symbols moved by one step


def test_container_v2_indexed_format(compressor, tmp_path):
    """the indexed container (magic LLZ2, streams back to back, 16-bit length table at the end) holds the
    SAME per-image rANS streams as the reference's v1 file, is 2 bytes per image smaller, and decompress_dataset reads
    either; a truncated / inconsistent v2 file raises."""
    comp, transform = compressor
    ds = _Synthetic(70, transform)
    f1, f2, lf = tmp_path / "Z1.bin", tmp_path / "Z2.bin", tmp_path / "Y.npy"
    comp.compress_dataset(ds, f1, label_file=lf, kwargs_dataloader=dict(batch_size=32, num_workers=0), is_info=False)
    comp.compress_dataset(ds, f2, label_file=lf, kwargs_dataloader=dict(batch_size=32, num_workers=0), is_info=False, container="v2")
    b1, b2 = f1.read_bytes(), f2.read_bytes()
    assert b2[:4] == b"LLZ2" and struct.unpack(">II", b2[4:12]) == (70, 512)
    assert len(b2) == len(b1) + 8 - 2 * 70
    # same streams: walk v1's records and compare with v2's payload + length table
    lens = np.frombuffer(b2[len(b2) - 140:], dtype="<u2").astype(np.int64) * 4
    pos1, pos2 = 4, 12
    for n in lens:
        assert struct.unpack(">I", b1[pos1: pos1 + 4])[0] == n and b1[pos1 + 4: pos1 + 4 + n] == b2[pos2: pos2 + n]
        pos1, pos2 = pos1 + 4 + n, pos2 + n
    Z1 = comp.decompress_dataset(f1, is_info=False)
    Z2, Y = comp.decompress_dataset(f2, label_file=lf, is_info=False)
    assert np.array_equal(Z1, Z2) and Y.tolist() == [i % 10 for i in range(70)]
    (tmp_path / "bad2.bin").write_bytes(b2[:-3])
    with pytest.raises(ValueError):
        comp.decompress_dataset(tmp_path / "bad2.bin")
    with pytest.raises(ValueError):
        comp.compress_dataset(ds, tmp_path / "x.bin", container="v3")
