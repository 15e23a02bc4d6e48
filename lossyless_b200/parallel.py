"""Multi-GPU plumbing for the two paths that shard, one process per GPU
under torchrun / torch.distributed:

  * encode (``compress_dataset``): images are independent units -> every rank codes a contiguous
    slice of the dataset with NO collective on the data path; the per-rank record blobs are
    gathered once at the end and rank 0 writes header + blobs in rank order, so the file is
    byte-identical to the single-GPU file (hub/compressor.py:192-196 format);
  * rate-distortion training: data parallel, gradient all-reduce on the (small) trainable
    bucket only (``learnable_compressors.allreduce_mean_``) plus the reference's all_gather of
    logit tiles inside the contrastive distortion (``helpers.GatherFromGpus``).
"""
import struct
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_range", "gather_record_blobs", "write_container", "compress_dataset_sharded"]


def shard_range(n, rank, world):
    """Contiguous [start, end) slice of n units for `rank`; the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_record_blobs(blob, n_local, group=None):
    """Gather every rank's (record bytes, record count) on rank 0 -> (list of bytes, list of counts)
    in rank order on rank 0, (None, None) elsewhere.  Works with gloo (CPU) and nccl (CUDA)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = torch.tensor([len(blob), n_local], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    sizes = [int(m[0]) for m in metas]
    counts = [int(m[1]) for m in metas]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(blob):
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    if rank != 0:
        return None, None
    return [bufs[r][: sizes[r]].cpu().numpy().tobytes() for r in range(world)], counts


def write_container(file, blobs, counts):
    """``>I N`` then the records of every rank in rank (= dataset) order."""
    with Path(file).open("wb") as f:
        f.write(struct.pack(">I", sum(counts)))
        for b in blobs:
            f.write(b)


@torch.no_grad()
def compress_dataset_sharded(compressor, dataset, file, label_file=None, batch_size=128, group=None):
    """``ClipCompressor.compress_dataset`` over all ranks of the default process group.  Returns
    (n_images, file_bytes) on rank 0, (n_images_local, None) elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    start, end = shard_range(len(dataset), rank, world)
    blob, labels = bytearray(), []

    def batches():
        for i0 in range(start, end, batch_size):
            items = [dataset[i] for i in range(i0, min(end, i0 + batch_size))]
            if label_file is not None:
                labels.extend(int(it[1]) for it in items)
            yield torch.stack([it[0] for it in items]).pin_memory()

    # the rank's slice goes through compress_dataset's own pipelined loop (H2D of the next pass under the kernels
    # of the current one, small batches grouped into encoder passes of up to 1512 images)
    for records, _ in compressor.compress_stream(batches(), coalesce=1512):
        blob += records.numpy().tobytes()
    blobs, counts = gather_record_blobs(bytes(blob), end - start, group)
    if label_file is not None:
        lab = [None] * world if rank == 0 else None
        dist.gather_object(labels, lab, dst=0, group=group)
    if rank != 0:
        return end - start, None
    write_container(file, blobs, counts)
    if label_file is not None:
        np.save(label_file, np.concatenate([np.asarray(l, dtype=np.uint16) for l in lab]), allow_pickle=False)
    return sum(counts), Path(file).stat().st_size
