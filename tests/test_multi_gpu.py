"""GPU, world_size 2 over NCCL (skipped on a 1-GPU box): the two paths that shard.
  * encode: every rank codes a contiguous slice with no collective on the data path; the merged
    container file is byte-identical to the single-GPU file;
  * training: the contrastive distortion's all_gather of logit tiles + the gradient bucket all-reduce
    reproduce the oracle's two-rank loss (lossyless/distortions.py:327-336 quirk included)."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Images(torch.utils.data.Dataset):
    def __init__(self, n):
        g = torch.Generator().manual_seed(0)
        self.x = torch.randn(n, 3, 224, 224, generator=g).half()

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], i % 7


def _worker(rank, world, port, tmp):
    import warnings

    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    import torch.distributed as dist

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import hubconf
    from lossyless_b200.distortions import ContrastiveDistortion
    from lossyless_b200.learnable_compressors import allreduce_mean_
    from lossyless_b200.parallel import compress_dataset_sharded, decompress_dataset_sharded
    from oracle import losses as olosses

    comp, _ = hubconf.clip_compressor_b005(device=f"cuda:{rank}")
    ds = _Images(45)
    compress_dataset_sharded(comp, ds, os.path.join(tmp, "sharded.bin"), label_file=os.path.join(tmp, "y.npy"), batch_size=16)
    if rank == 0:
        comp.compress_dataset(ds, os.path.join(tmp, "single.bin"), kwargs_dataloader=dict(batch_size=16, num_workers=0), is_info=False)
    dist.barrier()
    # sharded decode of that file: records split by the prefix sum of their lengths, no collective on the data path,
    # rows gathered on rank 0 == the single-GPU decode
    out = decompress_dataset_sharded(comp, os.path.join(tmp, "sharded.bin"), label_file=os.path.join(tmp, "y.npy"))
    if rank == 0:
        Z, Y = out
        Z1 = comp.decompress_dataset(os.path.join(tmp, "sharded.bin"), is_info=False)
        assert Z.shape == (45, 512) and np.array_equal(Z, Z1) and Y.tolist() == [i % 7 for i in range(45)]
    else:
        assert out is None

    # contrastive distortion with the real all_gather: same projector on both ranks, different data
    gen = torch.Generator().manual_seed(7)
    proj = dict(w1=torch.randn(128, 512, generator=gen) * 0.05, b1=torch.zeros(128), w2=torch.randn(256, 128, generator=gen) * 0.1, b2=torch.zeros(256))
    zs = [torch.randn(8, 512, generator=gen) for _ in range(4)]
    m = ContrastiveDistortion(temperature=0.01, is_train_temperature=True, effective_batch_size=5000,
                              project_kwargs=dict(mode="mlp", out_shape=0.5, in_shape=512), is_already_featurized=True)
    with torch.no_grad():
        l1, l2 = m.projector.module[0], m.projector.module[-1]
        l1.weight.copy_(proj["w1"]), l1.bias.zero_(), l2.weight.copy_(proj["w2"]), l2.bias.zero_()
    m = m.cuda()
    z, zp = zs[2 * rank].cuda().requires_grad_(True), zs[2 * rank + 1].cuda()
    d, logs, _ = m(z, zp, None, None)
    other = olosses.contrastive_local_logits(zs[2 * (1 - rank)], zs[2 * (1 - rank) + 1], proj)
    ref, _ = olosses.contrastive_distortion(zs[2 * rank], zs[2 * rank + 1], proj, rank=rank, other_tiles=[other],
                                            temperature=0.01, logit_scale=torch.log(torch.tensor(1 / 0.07)), effective_batch_size=5000)
    assert torch.allclose(d.detach().cpu(), ref, rtol=1e-4, atol=1e-5), (d, ref)
    assert logs["n_negatives"] == 2 * 16 - 1
    d.sum().backward()
    params = [p for p in m.parameters() if p.grad is not None]
    before = [p.grad.clone() for p in params]
    allreduce_mean_(params)
    g0 = [torch.zeros_like(b) for b in before]
    for b, g in zip(before, g0):
        t = b.clone()
        dist.all_reduce(t)
        g.copy_(t / world)
    assert all(torch.allclose(p.grad, g) for p, g in zip(params, g0))

    # unfrozen CLIP tower (2 blocks to keep it short), data parallel: different data per rank, bf16 gradient
    # buckets; after one fit() step every rank holds the same parameters, and they moved
    from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams

    hp = default_clip_hparams(distortion="BINCE", beta=1e-2, max_steps=100, n_data=5000, is_frozen=False)
    torch.manual_seed(0)
    model = LearnableCompressor(hp)
    enc = model.p_ZlX.mapper.encoder
    del enc.transformer.resblocks[2:]
    enc.layers = 2
    model = model.cuda().train()
    w0 = enc.transformer.resblocks[1].mlp.c_fc.weight.detach().clone()
    g = torch.Generator().manual_seed(50 + rank)
    batch = (torch.randn(4, 3, 224, 224, generator=g).cuda(), (None, torch.randn(4, 3, 224, 224, generator=g).cuda()))
    torch.manual_seed(9 + rank)
    model.fit([batch], grad_comm_dtype=torch.bfloat16)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    other_flat = flat.clone()
    dist.broadcast(other_flat, src=0)
    assert torch.equal(flat, other_flat), "ranks diverged after the gradient all-reduce"
    assert (enc.transformer.resblocks[1].mlp.c_fc.weight.detach() - w0).abs().max().item() > 0
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
        a = open(os.path.join(tmp, "sharded.bin"), "rb").read()
        b = open(os.path.join(tmp, "single.bin"), "rb").read()
        y = np.load(os.path.join(tmp, "y.npy"))
    assert a == b
    assert y.tolist() == [i % 7 for i in range(45)] and y.dtype == np.uint16


def _worker_w8(rank, world, port):
    """SURVEY 8(d) configuration 3's world size over real NCCL: every rank feeds the product ContrastiveDistortion the
    inputs the reference's own eight-process gloo run used (tests/golden/ref_losses_w8.npz); ranks 0 / 3 / 7 hold the
    reference's logits, distortion, negatives count and gradients."""
    import warnings

    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    import torch.distributed as dist

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from lossyless_b200.distortions import ContrastiveDistortion

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_losses.npz"))
    g8 = np.load(os.path.join(ROOT, "tests", "golden", "ref_losses_w8.npz"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from ref_configs import CONTRASTIVE_KW

    m = ContrastiveDistortion(**{**CONTRASTIVE_KW, "project_kwargs": dict(CONTRASTIVE_KW["project_kwargs"])})
    m.load_state_dict({k[len("contrastive/"):]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("contrastive/")})
    m = m.cuda()
    z = torch.from_numpy(g8[f"w8/rank{rank}/z_hat"]).cuda().requires_grad_(True)
    zp = torch.from_numpy(g8[f"w8/rank{rank}/z_pos"]).cuda().requires_grad_(True)
    logits = m.compute_logits(z, zp, None)
    d, logs, _ = m(z, zp, None, None)
    d.mean().backward()
    assert logits.shape == (8, 64) and int(logs["n_negatives"]) == 63
    if rank in (0, 3, 7):
        assert torch.allclose(logits.detach().cpu(), torch.from_numpy(g8[f"w8/rank{rank}/logits"]), atol=2e-6)
        assert np.allclose(d.detach().cpu().numpy(), g8[f"w8/rank{rank}/distortion"], rtol=1e-4)
        for got, want in ((z.grad, "g_z_hat"), (zp.grad, "g_z_pos"), (m.logit_scale.grad, "g_logit_scale")):
            w = torch.from_numpy(np.asarray(g8[f"w8/rank{rank}/{want}"]))
            assert (got.detach().cpu() - w).abs().max() <= 2e-3 * w.abs().max() + 1e-9, (rank, want)
    dist.barrier()
    dist.destroy_process_group()


def test_eight_gpus_contrastive_matches_reference_run(cuda_device):
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs (gpurun --gpus 8)")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_w8, args=(8, port), nprocs=8, join=True)
