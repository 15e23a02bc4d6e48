"""Multi-GPU plumbing for the two paths that shard, one process per GPU
under torchrun / torch.distributed:

  * encode (``compress_dataset``): images are independent units -> every rank codes a contiguous
    slice of the dataset with NO collective on the data path; the per-rank record blobs are
    gathered once at the end and rank 0 writes header + blobs in rank order, so the file is
    byte-identical to the single-GPU file (hub/compressor.py:192-196 format);
  * decode (``decompress_dataset``): records are independent too -> the container is indexed on every rank (a
    host-side walk over the length prefixes), the records are split into contiguous ranges by the PREFIX SUM OF
    THEIR LENGTHS (equal bytes per rank: a stream's decode time follows its length), every rank decodes its range
    with no collective on the data path, and the rows are gathered once on rank 0 in record order
    (hub/compressor.py:209-254 is the serial loop this replaces);
  * rate-distortion training: data parallel, gradient all-reduce on the (small) trainable
    bucket only (``learnable_compressors.allreduce_mean_``) plus the reference's all_gather of
    logit tiles inside the contrastive distortion (``helpers.GatherFromGpus``).
"""
import struct
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_range", "gather_record_blobs", "write_container", "compress_dataset_sharded",
           "split_records_by_bytes", "gather_rows", "decompress_dataset_sharded"]


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index, sysfs="/sys/bus/pci/devices"):
    """Pin this process (one rank per GPU) to the CPUs of the NUMA node its GPU hangs off, and prefer that node for its
    memory, BEFORE it allocates pinned staging buffers: with eight ranks feeding eight GPUs, staging memory that
    first-touch placed on the other socket crosses the inter-socket link on every host-to-device copy (round 1
    measured 24 GB/s per GPU with all eight copying against 55 GB/s alone).  Pure sysfs + sched_setaffinity +
    set_mempolicy(MPOL_PREFERRED); returns ``{"node", "cpus", "pci"}`` or ``None`` when the topology is not exposed
    (containers without sysfs PCI entries, single-node hosts): then nothing is changed."""
    import ctypes
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        pci = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = os.path.join(sysfs, pci)
        node = int(open(os.path.join(base, "numa_node")).read())
        cpus = _parse_cpulist(open(os.path.join(base, "local_cpulist")).read())
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
    if node < 0 or not cpus:
        return None
    allowed = os.sched_getaffinity(0) & cpus
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    try:  # MPOL_PREFERRED = 1: allocate on `node` when it has room, fall back elsewhere instead of failing
        import platform

        nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())  # __NR_set_mempolicy
        if nr is not None and node < 8 * ctypes.sizeof(ctypes.c_ulong):
            mask = ctypes.c_ulong(1 << node)
            ctypes.CDLL(None, use_errno=True).syscall(nr, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
    except (OSError, AttributeError):
        pass
    return {"node": node, "cpus": len(allowed), "pci": pci}


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


def split_records_by_bytes(lens, world):
    """Contiguous record ranges [(lo, hi)] * world with (nearly) equal payload bytes: rank r takes the records whose
    cumulative length falls in [r, r + 1) * total / world.  Every record lands in exactly one range, in order."""
    lens = np.asarray(lens, dtype=np.int64)
    n = len(lens)
    if n == 0:
        return [(0, 0)] * world
    csum = np.cumsum(lens + 4)  # + the 4-byte length prefix each record carries
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, r * total / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def gather_rows(rows, counts, group=None):
    """Row blocks of every rank (``rows`` [counts[rank], C] float32, device tensor under nccl, CPU tensor under gloo)
    -> on rank 0 one [sum(counts), C] CPU tensor in rank order; None elsewhere.  One padded gather."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = max(max(counts), 1)
    C = rows.shape[1]
    pad = torch.zeros((cap, C), dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([bufs[r][: counts[r]].cpu() for r in range(world)], dim=0)


@torch.no_grad()
def decompress_dataset_sharded(compressor, file, label_file=None, group=None, decode_fn=None):
    """``ClipCompressor.decompress_dataset`` over all ranks of the process group.  Returns ``Z_hat`` float32
    [N, 512] (and ``Y``) on rank 0, None elsewhere.  ``decode_fn(data, offsets, lens, lo, hi) -> [hi - lo, C]
    tensor`` replaces the compressor's GPU decode (the gloo test injects a host decoder)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    data, offsets, lens = compressor.load_container(file)
    spans = split_records_by_bytes(lens, world)
    lo, hi = spans[rank]
    if decode_fn is not None:
        z = decode_fn(data, offsets, lens, lo, hi)
    else:
        dev = torch.device(compressor.device)
        z = torch.empty((hi - lo, compressor.z_dim), dtype=torch.float32, device=dev)
        compressor.decode_records(data, offsets, lens, lo, hi, out=z)
    Z = gather_rows(z, [b - a for a, b in spans], group)
    if rank != 0:
        return None
    Z = Z.numpy()
    if label_file is not None:
        return Z, np.load(label_file, allow_pickle=False).astype(np.int64)
    return Z
