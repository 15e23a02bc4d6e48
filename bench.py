#!/usr/bin/env python
"""Benchmark of the encode hot path: img/sec through CLIP ViT-B/32 + EntropyBottleneck + rANS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode encode|train]

One "step" = one batch of synthetic CLIP-normalised 3x224x224 fp16 images through
``clip_compressor_b005`` (ViT-B/32 -> process_z_in -> quantise -> rANS -> container records),
i.e. the body of the reference's ``compress_dataset`` loop (hub/compressor.py:186-189).
Prints ONE JSON line on rank 0 (contract in the task statement):
  value     images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through ``ClipCompressor.compress_stream`` (compress_dataset's loop) with pinned HOST
            buffers: H2D of the images and D2H of the records inside the timed region.  The images are CLIP-sized
            uint8 pixels (ToTensor / Normalize on the GPU, same container bytes: tests/test_preprocess_gpu.py);
            ``e2e_fp16_feed`` is the same with pre-normalised fp16 tensors (twice the bytes over the host link)
  roofline  the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / event-timed GEMM time,
            against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the same path on the host cores, bounded sample (N=1)
  parity    bits/img of BOTH arms on the SAME images and the SAME (fp16-rounded) ViT weights for the three
            published bottlenecks: GPU vs fp32 CPU oracle, symbol-flip rate of the 16-bit encoder, and the
            byte-identity of the EB + rANS stage given identical z (N=1)
  decode    decompress_dataset: device-resident and end-to-end (file -> Z_hat in host memory) img/s, CPU oracle beside
With ``--impl reference`` the CPU restatement of the reference's own call pattern is timed
instead (the reference itself cannot be installed: compressai / clip are not available), on the same
images and weights, 128 images per step (the reference's DataLoader batch, hub/compressor.py:155), with every
host core (torchrun exports OMP_NUM_THREADS=1: it is undone here).
``--mode train`` times BASELINE config 3 instead: the rate-distortion training step (CLIP ViT-B/32 unfrozen,
BINCE, beta = 0.01, per-GPU batch 256, bf16 gradient buckets over NCCL) -- the one path with collectives.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_GEMM_PER_IMG = 2 * (49 * 3072 * 768 + 12 * 50 * (768 * 2304 + 768 * 768 + 2 * 768 * 3072) + 768 * 512) / 1e9
# the last block's out-projection / MLP on the 49 non-class-token rows: computed by the reference, never read by it, not
# executed here (ll_vit_set_cls_last_block / clip_vit.CLS_LAST_BLOCK) -- the rooflines count EXECUTED FLOPs only
GFLOP_SKIPPED_PER_IMG = 2 * 49 * (768 * 768 + 2 * 768 * 3072) / 1e9
GFLOP_EXECUTED_PER_IMG = GFLOP_GEMM_PER_IMG - GFLOP_SKIPPED_PER_IMG
METRIC = "img/sec encode (CLIP+entropy+rANS) 224x224"
WORKLOAD = "clip_compressor_b005 compress_dataset body on synthetic CLIP-normalised 3x224x224 fp16 batches"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU sampled while the timed region runs: NVML queried
    every 10 ms from a thread (the timed region of the default run is ~150 ms, and with eight ranks an
    `nvidia-smi -lms` child per rank does not even deliver its first line in that time); falls back to the
    nvidia-smi loop of the profiling recipe if NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop = index, [], None, None, False
        self.samples, self.max_mhz, self.reason_bits = [], None, 0

    def _nvml_handle(self):
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES may renumber the devices: look the GPU up by UUID
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # raises here if the query is unsupported
        return pynvml, h

    def __enter__(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(reasons(self.handle))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.nvml is not None:
            self.stop = True
            self.thread.join(timeout=2)
        elif self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        if self.nvml is not None and self.samples:
            return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(k for k, b in self.BITS.items() if self.reason_bits & b), "samples": len(self.samples),
                    "source": "nvml, 10 ms"}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 200"}


# ----------------------------------------------------------------------------- shared inputs
def bench_images(n, rank=0):
    """The synthetic batch BOTH arms use: CLIP-normalised-like N(0,1) fp16 images from a CPU generator (so the
    CPU reference arm and the GPU arm see the same pixels), seed 1234 + rank."""
    import torch

    return torch.randn(n, 3, 224, 224, generator=torch.Generator().manual_seed(1234 + rank)).half()


def bench_vit_state_dict():
    """The ViT-B/32 weights BOTH arms use: the package's seeded random init (no pretrained CLIP offline), rounded
    to fp16 as the hub model stores them, as float32 CPU tensors under CLIP's own key names."""
    import torch

    from lossyless_b200.clip_vit import VisionTransformer

    vit = VisionTransformer(224, 32, 768, 12, 12, 512)
    return {k: v.detach().half().float() for k, v in vit.state_dict().items()}


# ----------------------------------------------------------------------------- CPU legs
def cpu_reference_pipeline(x, faithful=True, tag="beta5e-02"):
    """The reference's path restated on the CPU (oracle): fp32 ViT-B/32 (torch CPU, all host
    threads) -> process_z_in -> quantise -> one rANS call per image.  faithful=True mirrors
    compressai 1.1.5's call pattern (per-image Python loop, `.tolist()` of the symbols and of the
    whole CDF table, hub/compressor.py:98).  x: fp16 images.  Returns (seconds, bits_per_img, z fp16, strings)."""
    import numpy as np
    import torch

    from oracle import eb as oeb
    from oracle import rans
    from oracle import vit as ovit

    st = _cpu_state(tag)
    n_images = x.shape[0]
    t0 = time.perf_counter()
    with torch.no_grad():
        z = ovit.forward(_cpu_vit(), x.float()).half()  # reference: fp16 z
    z_in = oeb.process_z_in(z, st["sd"]["biasing"], st["sd"]["scaling"])
    sym = torch.round(z_in - oeb.medians(st["eb"])[None]).int()
    tab = st["tab"]
    if faithful:
        idx = list(range(512))
        cdf_t, len_t, off_t = torch.from_numpy(tab["cdf"]), torch.from_numpy(tab["cdf_length"]), torch.from_numpy(tab["offset"])
        strings = []
        for i in range(n_images):
            strings.append(rans.encode_with_indexes(sym[i].reshape(-1).int().tolist(), idx, np.array(cdf_t.tolist(), dtype=np.int32),
                                                    len_t.reshape(-1).int().tolist(), off_t.reshape(-1).int().tolist()))
    else:
        strings = rans.encode_batch(sym.numpy(), tab["cdf"], tab["cdf_length"], tab["offset"])
    nbytes = sum(len(s) + 4 for s in strings)
    return time.perf_counter() - t0, 8.0 * nbytes / n_images, z, strings


_CPU = {}


def _cpu_vit():
    if "vit" not in _CPU:
        _CPU["vit"] = bench_vit_state_dict()
    return _CPU["vit"]


def _cpu_state(tag="beta5e-02"):
    if tag not in _CPU:
        import torch

        from oracle import eb as oeb

        sd = torch.load(os.path.join(ROOT, "lossyless_b200", "hub", "weights", f"{tag}_factorized_rate.pt"),
                        map_location="cpu", weights_only=True)
        eb = oeb.split_state_dict(sd)
        _CPU[tag] = dict(sd=sd, eb=eb, tab=oeb.update(eb))
    return _CPU[tag]


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is ONE process and is entitled to the box."""
    import torch

    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


REF_BATCH = 128  # hub/compressor.py:155 kwargs_dataloader=dict(batch_size=128, ...)


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path (oracle port; compressai / clip cannot be
    installed) timed on this box's host cores; rank 0 only (the other ranks exit 0 without work)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    cores = use_all_host_threads()
    n = args.cpu_images or REF_BATCH
    X = bench_images(max(n, min(1512, n * max(args.steps, 1))))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pipeline(X[:4])
    t, bits = 0.0, []
    for s in range(args.steps):
        lo = (s * n) % max(1, X.shape[0] - n + 1)
        dt, b, _, _ = cpu_reference_pipeline(X[lo: lo + n])
        t += dt
        bits.append(b)
    value = n * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n} images per step on the host CPU (the reference's DataLoader batch)",
                   "bits_per_img": sum(bits) / len(bits), "inputs": "bench_images(seed 1234): the GPU arm's own images and ViT weights"},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {n} images: fp32 torch-CPU ViT-B/32 + per-image rANS "
                                   f"(compressai call pattern), torch threads {cores}, os.cpu_count()={os.cpu_count()}, "
                                   f"OMP_NUM_THREADS env was {os.environ.get('OMP_NUM_THREADS', 'unset')}"},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def parity_leg(comp, X, z_gpu, n=256):
    """Both arms on the SAME n images and the SAME ViT weights, for the three published bottlenecks: bits/img of the
    GPU path (fp16 tcgen05 ViT -> fused quantise + rANS kernel) against the fp32 CPU oracle, the symbol-flip rate
    the 16-bit encoder causes, and whether the EB + rANS stage is byte-identical when both are fed the GPU's z."""
    import numpy as np
    import torch

    from lossyless_b200.rates import get_rate_estimator
    from oracle import eb as oeb
    from oracle import rans
    from oracle import vit as ovit

    n = min(n, X.shape[0])
    x_cpu = X[:n].cpu()
    t0 = time.perf_counter()
    with torch.no_grad():
        z_f32 = ovit.forward(_cpu_vit(), x_cpu.float())
    t_vit = time.perf_counter() - t0
    z_ref = z_f32.half()  # the reference's clip(X) is fp16
    zg = z_gpu[:n]
    dz = (zg.float().cpu() - z_f32).abs()
    out = {"images": n, "weights": "identical: the GPU arm's fp16 ViT-B/32 weights, run in fp32 by the oracle",
           "vit_abs_err_vs_fp32": {"max": float(dz.max()), "mean": float(dz.mean()), "z_rms": float(z_f32.pow(2).mean().sqrt())},
           "oracle_vit_img_per_s": n / t_vit}
    for tag, beta in (("beta1e-01", 0.1), ("beta5e-02", 0.05), ("beta1e-02", 0.01)):
        st = _cpu_state(tag)
        r = get_rate_estimator("H_factorized", z_dim=512, kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))
        r.load_state_dict(st["sd"])
        r.to(zg.device).eval()
        r.prepare_compressor_()
        with torch.no_grad():
            s_gpu = r.compress(zg)[0]
        med = oeb.medians(st["eb"])[None]
        sym_ref = torch.round(oeb.process_z_in(z_ref, st["sd"]["biasing"], st["sd"]["scaling"]) - med).int()
        sym_gpu = torch.round(oeb.process_z_in(zg.cpu(), st["sd"]["biasing"], st["sd"]["scaling"]) - med).int()
        tab = st["tab"]
        s_ref = rans.encode_batch(sym_ref.numpy(), tab["cdf"], tab["cdf_length"], tab["offset"])
        s_same_z = rans.encode_batch(sym_gpu.numpy(), tab["cdf"], tab["cdf_length"], tab["offset"])
        bits_gpu = 8.0 * sum(len(s) + 4 for s in s_gpu) / n
        bits_ref = 8.0 * sum(len(s) + 4 for s in s_ref) / n
        out[tag] = {"beta": beta, "bits_per_img_gpu": bits_gpu, "bits_per_img_oracle": bits_ref,
                    "delta_bits_pct": 100.0 * (bits_gpu - bits_ref) / bits_ref,
                    "symbol_flip_rate": float((sym_gpu != sym_ref).float().mean()),
                    "max_symbol_delta": int((sym_gpu - sym_ref).abs().max()),
                    "bytes_identical_given_z": s_gpu == s_same_z}
    return out


# ----------------------------------------------------------------------------- GPU arm
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (average over one block's four GEMMs) from the
    committed ncu --set full capture; None if the summary is not in the tree."""
    prof = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    for name in ("r02_ncu_vit_gemms_dram_traffic_final.json", "r02_ncu_vit_gemms_dram_traffic.json", "r01_ncu_vit_gemms_dram_traffic.json"):
        try:
            with open(os.path.join(prof, name)) as f:
                return float(json.load(f)["per_launch_dram_bytes_layer_average"]), name
        except (OSError, KeyError, ValueError):
            continue
    return None, None


NUMA_BINDING = None


def dist_setup():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the measured path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host side of the feed: this rank's threads and pinned staging memory on the GPU's own NUMA node (LL_NO_NUMA=1: A/B)
    global NUMA_BINDING
    if os.environ.get("LL_NO_NUMA", "0") != "1":
        from lossyless_b200.parallel import bind_host_to_gpu

        NUMA_BINDING = bind_host_to_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    return rank, local_rank, world, dev, barrier, (lambda v: reduce(v, dist.ReduceOp.MAX)), (lambda v: reduce(v, dist.ReduceOp.SUM))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="encode", choices=["encode", "train"])
    ap.add_argument("--batch", type=int, default=0, help="units per step per GPU (encode: 1512 images = one encoder pass of 4 full rounds of 74 CTA pairs x 256 token rows; train: 256 samples)")
    ap.add_argument("--chunk", type=int, default=0, help="images per internal ViT pass (0 = library default)")
    ap.add_argument("--cpu-images", type=int, default=0, help="images per CPU-baseline step (0: 128 for --impl reference, 32 for the cpu_baseline leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse-ln", action="store_true", help="A/B: run ln_1 / ln_2 as separate kernels")
    ap.add_argument("--attn-impl", type=int, default=-1, help="A/B: ll_attention_set_impl (0 cp.async, 1 tcgen05, 2 TMA); -1 = library default")
    ap.add_argument("--roofline-seconds", type=float, default=2.0, help="length of the event-bracketed roofline leg")
    ap.add_argument("--frozen", action="store_true", help="--mode train: freeze the CLIP tower (config/finetune/freezer.yaml)")
    ap.add_argument("--distortion", default="BINCE", choices=["BINCE", "lossy_Z"], help="--mode train")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)
    warnings.filterwarnings("ignore", message="no CLIP weights given")
    if args.mode == "train":
        return run_train(args)
    return run_encode(args)


def run_encode(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world, dev, barrier, max_over_ranks, sum_over_ranks = dist_setup()
    import hubconf

    comp, _ = hubconf.clip_compressor_b005(device=f"cuda:{local_rank}")
    if args.chunk:
        comp.clip.set_chunk(args.chunk)
    if args.no_fuse_ln:
        comp.clip.set_fuse_layernorm(False)
    if args.attn_impl >= 0:
        from lossyless_b200 import _native as _N

        _N.call("ll_attention_set_impl", args.attn_impl)
    B = args.batch or 1512
    # encode shards by image: every rank codes its own independent batch, no collective on the path
    X_host = bench_images(B, rank).pin_memory()
    X = X_host.to(dev)

    def step_device():
        return comp._compress_records(X, mode=1)

    for _ in range(args.warmup):
        packed, offsets, lens = step_device()
    barrier()
    if not bool(torch.isfinite(comp.clip(X[:64]).float()).all()):
        raise SystemExit("bench.py: the encoder produced non-finite representations; refusing to time a broken path")
    bits_per_img = 8.0 * (float(offsets[-1].item()) / B)  # incl. the 4-byte length prefix, as README's rate
    total_bytes = int(offsets[-1].item())

    # ---- value: device-resident inputs, CUDA events on the launch stream
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_device()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: pinned host images in, container records out, per step, through compress_dataset's own loop
    # (ClipCompressor.compress_stream: batch i+1 is enqueued before the host waits for batch i's records; every
    # step's H2D copy and records D2H happen inside the timed region)
    rec_bytes = [0]

    def e2e_run(x_host, steps):
        n = 0
        for rec, b in comp.compress_stream(x_host for _ in range(steps)):
            n += b
            rec_bytes[0] = int(rec.numel())  # bytes of the last batch's records that came back to the host
        return n

    e2e_run(X_host, 2)
    barrier()
    t0 = time.perf_counter()
    n_done = e2e_run(X_host, args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert n_done == B * args.steps
    e2e = world * B * args.steps / e2e_s
    # the same, one blocking call per batch (no overlap across batches)
    pinned_out = None
    _, _, pinned_out = comp.compress_batch_to_host(X_host, pinned_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rec, _, pinned_out = comp.compress_batch_to_host(X_host, pinned_out)
    barrier()
    e2e_blocking = world * B * args.steps / max_over_ranks(time.perf_counter() - t0)

    # ---- e2e from raw pixels: STL10-shaped uint8 images (BASELINE config 2) -> H2D of the bytes -> GPU CLIP
    # preprocessing (PIL-exact bicubic 96 -> 224, normalise) -> ViT -> codec -> records D2H
    raw_host = torch.randint(0, 256, (B, 96, 96, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(rank)).pin_memory()
    e2e_run(raw_host, 2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(raw_host, args.steps)
    barrier()
    e2e_raw = world * B * args.steps / max_over_ranks(time.perf_counter() - t0)
    # ---- e2e from CLIP-sized uint8 pixels (224 x 224 x 3, 150 KB per image instead of 301 KB of fp16): what a
    # DataLoader that resizes on the CPU but leaves ToTensor / Normalize to the GPU feeds
    u8_host = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
    e2e_run(u8_host, 2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(u8_host, args.steps)
    barrier()
    e2e_u8 = world * B * args.steps / max_over_ranks(time.perf_counter() - t0)
    u8_rec_bytes = rec_bytes[0]

    # ---- what the host link gives: a bare pinned H2D copy of one 378-image sub-batch, nothing else running
    probe_dst = torch.empty_like(X[:378])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_dst.copy_(X_host[:378], non_blocking=True)
    ev0.record()
    for _ in range(4):
        probe_dst.copy_(X_host[:378], non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbps = 4 * probe_dst.numel() * 2 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    h2d_min = -max_over_ranks(-h2d_gbps)

    # ---- decode: decompress_dataset.  value = kernels only, streams resident in HBM; e2e = the public call on a
    # container file (file read into pinned memory, index, H2D, decode, Z_hat back in host memory)
    import tempfile

    reps = 8
    blob = bytes(packed[:total_bytes].cpu().numpy().data)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, f"Z{rank}.bin")
        with open(path, "wb") as f:
            f.write((reps * B).to_bytes(4, "big"))
            for _ in range(reps):
                f.write(blob)
        data_h, off_h, lens_h = comp.load_container(path)
        n_dec = len(lens_h)
        z_dev = torch.empty((n_dec, 512), dtype=torch.float32, device=dev)
        comp.decode_records(data_h, off_h, lens_h, out=z_dev)
        data_d = data_h.to(dev)
        off_d, lens_d = torch.from_numpy(off_h).to(dev), torch.from_numpy(lens_h).to(dev)
        status = torch.zeros(n_dec, dtype=torch.int32, device=dev)
        es = comp._exp_scale()

        def decode_device():
            comp.entropy_bottleneck.decompress_from_device(data_d, off_d, lens_d, bias=comp.biasing, exp_scale=es, status_out=status, out=z_dev)

        decode_device()
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            decode_device()
        d1.record()
        barrier()
        dec_ms = max_over_ranks(d0.elapsed_time(d1))
        assert int(status.sum().item()) == 0
        comp.decompress_dataset(path, is_info=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            Zd = comp.decompress_dataset(path, is_info=False)
        barrier()
        dec_e2e_s = max_over_ranks(time.perf_counter() - t0) / 3
        assert Zd.shape == (n_dec, 512)
    decode = {"value": world * n_dec * args.steps / (dec_ms / 1e3), "unit": "img/s", "images": n_dec, "ms_per_call": dec_ms / args.steps,
              "bytes_per_img": {"in": float(lens_h.mean()), "out": 2048},
              "achieved_gbs": (float(lens_h.sum()) + 2048.0 * n_dec) * args.steps / (dec_ms / 1e3) / 1e9,
              "e2e": {"value": world * n_dec / dec_e2e_s, "unit": "img/s", "api": "ClipCompressor.decompress_dataset(file) -> np.ndarray",
                      "h2d_bytes_per_call": int(off_h[-1]), "d2h_bytes_per_call": n_dec * 2048}}

    # ---- roofline leg: the same steps with every GEMM launch bracketed by events, for >= --roofline-seconds
    n_prof = max(5, int(args.roofline_seconds / max(ms / args.steps / 1e3, 1e-4)))
    n_prof = min(n_prof, 400)
    comp.clip.profile_begin()
    for _ in range(n_prof):
        step_device()
    gemm_ms, gemm_flops, gemm_launches = comp.clip.profile_end()
    breakdown = {k: v / n_prof for k, v in comp.clip.profile_breakdown().items()}
    pk = peaks()
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    vit_launches = comp.clip.launches(B)
    traffic, traffic_src = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 (fp32 accumulate; integer rANS)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "beta": 0.05, "vit_weights": "seeded random init (no pretrained CLIP offline)",
                   "l2": f"inputs {B * 3 * 224 * 224 * 2 / 1e6:.0f} MB/step > 126 MB L2, no flush needed",
                   "bits_per_img": bits_per_img, "sharding": "independent images per rank, no collective"},
        # headline end-to-end number: the CLIP-sized uint8 feed.  The fp16 feed of round 1 (301 KB per image) is bound by
        # the host side of the link once eight ranks copy at once (24 GB/s per GPU on this pool's 8-GPU boxes: a VM with
        # one NUMA node, nothing to bind to); 150 KB of pixels per image, ToTensor / Normalize on the GPU inside the
        # timed region, is the same container byte for byte and keeps the link under the kernels at every N.
        "e2e": {"value": e2e_u8, "unit": "img/s", "h2d_bytes_per_step": B * 224 * 224 * 3, "d2h_bytes_per_step": u8_rec_bytes + 8,
                "input": "pinned host uint8 [B,224,224,3] CLIP-sized pixels (what a DataLoader holds after decode + resize); ToTensor / Normalize "
                         "on the GPU inside the timed region, then ViT -> bottleneck -> rANS -> records back in pinned host memory",
                "api": "ClipCompressor.compress_stream (compress_dataset's loop)", "host_numa_binding": NUMA_BINDING,
                "h2d_link_gbps": h2d_gbps, "h2d_link_gbps_min_over_ranks": h2d_min},
        "e2e_fp16_feed": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": B * 3 * 224 * 224 * 2, "d2h_bytes_per_step": total_bytes + 8,
                          "input": "pinned host fp16 [B,3,224,224] CLIP-normalised tensors (round 1's e2e: what the reference's loop body receives)",
                          "blocking_per_batch_call": e2e_blocking},
        "e2e_raw_uint8": {"value": e2e_raw, "unit": "img/s", "h2d_bytes_per_step": B * 96 * 96 * 3,
                          "input": "pinned host uint8 [B,96,96,3] STL10-shaped images; resize/crop/normalise on the GPU"},
        "decode": decode,
        "gpu_launches": args.steps * (vit_launches + 1 + 2),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all ViT linears)", "achieved": achieved,
                     "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
                     "traffic_unit": f"bytes of DRAM read + write per GEMM launch (ncu --set full capture, profiles/{traffic_src})",
                     "peak_source": pk["source"], "gemm_ms_per_step": gemm_ms / n_prof,
                     "gemm_launches_per_step": gemm_launches // n_prof, "leg_steps": n_prof, "leg_seconds": n_prof * ms / args.steps / 1e3,
                     "algorithmic_gflop_per_img": GFLOP_EXECUTED_PER_IMG, "reference_gflop_per_img": GFLOP_GEMM_PER_IMG,
                     "kernel_ms_per_step": breakdown},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = use_all_host_threads()
        n = args.cpu_images or 32
        cpu_reference_pipeline(X_host[:4])
        t, _, _, _ = cpu_reference_pipeline(X_host[:n])
        reps_c = max(1, min(8, int(10.0 / max(t, 1e-3))))
        tt = t + sum(cpu_reference_pipeline(X_host[(i + 1) * n: (i + 2) * n])[0] for i in range(reps_c - 1))
        line["cpu_baseline"] = {"value": n * reps_c / tt, "unit": "img/s", "cores": cores, "kind": "port",
                                "sample": f"{reps_c} x {n} of the step's own images: fp32 torch-CPU ViT-B/32 + per-image rANS oracle (compressai call pattern), os.cpu_count()={os.cpu_count()}"}
        with torch.no_grad():
            z_gpu = comp.clip(X[:256])
        line["parity"] = parity_leg(comp, X, z_gpu, 256)
        # CPU decode beside the GPU decode: the oracle's C decoder, single thread, on a bounded sample of the same file
        from oracle import eb as oeb

        st = _cpu_state("beta5e-02")
        raw = data_h.numpy()
        n_cpu = min(2048, n_dec)
        strings = [raw[off_h[i]: off_h[i] + lens_h[i]].tobytes() for i in range(n_cpu)]
        t0 = time.perf_counter()
        zd, _ = oeb.decompress(st["eb"], st["tab"], strings)
        zd = oeb.process_z_out(zd, st["sd"]["biasing"], st["sd"]["scaling"])
        t_dec = time.perf_counter() - t0
        line["decode"]["cpu_baseline"] = {"value": n_cpu / t_dec, "unit": "img/s", "cores": 1, "kind": "port",
                                          "sample": f"{n_cpu} records of the same file, batched C rANS decoder (oracle), one thread",
                                          "matches_gpu": bool(torch.equal(zd, z_dev[:n_cpu].cpu()))}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- training mode (BASELINE config 3)
TRAIN_METRIC = "img/sec rate-distortion training step (CLIP ViT-B/32, H_factorized + contrastive, beta=0.01)"


def run_train(args):
    """One step = LearnableCompressor.fit's loop body on one per-GPU batch: forward of both views through the CLIP
    tower, EntropyBottleneck rate, contrastive (BINCE) distortion incl. the logits-tile all_gather, backward through
    the tower, bf16 gradient-bucket all-reduce over NCCL, clip, AdamW ("main") + the coder optimiser pass."""
    import torch
    import torch.distributed as dist

    from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams

    rank, local_rank, world, dev, barrier, max_over_ranks, sum_over_ranks = dist_setup()
    B = args.batch or 256
    torch.manual_seed(1234)  # identical parameters on every rank (the ViT init is seeded separately)
    hp = default_clip_hparams(distortion=args.distortion, beta=1e-2, max_steps=1000, n_data=50000, is_frozen=args.frozen)
    model = LearnableCompressor(hp).to(dev).train()
    fit_kw = {} if args.frozen else dict(grad_comm_dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    xp_host = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    x, x_pos = x_host.to(dev), xp_host.to(dev)
    batch = (x, (None, x_pos))
    model.fit([batch] * args.warmup, **fit_kw)
    barrier()
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        logged = dict(model.fit([batch] * args.steps, **fit_kw))  # a copy: later fit() calls update the same dict
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms / 1e3)

    # the same steps without the gradient all-reduce (ranks drift apart: timing only) -> exposed collective time
    exposed = None
    if world > 1:
        model.fit([batch] * 2, skip_grad_sync=True, **fit_kw)
        barrier()
        e0.record()
        model.fit([batch] * args.steps, skip_grad_sync=True, **fit_kw)
        e1.record()
        barrier()
        ms_nosync = max_over_ranks(e0.elapsed_time(e1))
        exposed = (ms - ms_nosync) / args.steps

    # e2e: host batches (pinned fp32, what a DataLoader hands over) -> H2D inside the step, loss read back per step
    def e2e_steps(n):
        # fit() takes the HOST batches themselves: it copies batch i + 1 on a side stream under the kernels of batch i
        # and sends every step's loss to pinned host memory (model.loss_history) without stalling the host
        lg = model.fit(((x_host, (None, xp_host)) for _ in range(n)), **fit_kw)
        assert model.loss_history.numel() == n and bool(torch.isfinite(model.loss_history).all())
        return lg["train/feat/loss"]

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e = world * B * args.steps / max_over_ranks(time.perf_counter() - t0)
    views = 2 if args.distortion == "BINCE" else 1
    passes = 1 if args.frozen else 3  # forward (+ backward = 2x) through the tower
    cls_last = os.environ.get("LL_CLS_LAST_BLOCK", "1") != "0"
    gflop_img = GFLOP_EXECUTED_PER_IMG if cls_last else GFLOP_GEMM_PER_IMG
    flops = world * B * views * passes * gflop_img * 1e9
    pk = peaks()
    achieved = flops / (ms / args.steps / 1e3) / 1e12 / world
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    line = {
        "metric": TRAIN_METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 GEMMs (fp32 accumulate, fp32 master weights, fp32 entropy model)" if not args.frozen else "f16 tower, fp32 entropy model",
        "data": "synthetic",
        "config": {"workload": f"RD training step, CLIP ViT-B/32 {'frozen' if args.frozen else 'unfrozen'}, rate=H_factorized, distortion={args.distortion}, "
                               f"per-GPU batch {B}, AdamW 1e-3 / 3e-4, clip 3 (SURVEY 8d config 3)",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}", "trainable_params": n_params,
                   "l2": f"inputs {2 * B * 3 * 224 * 224 * 4 / 1e6:.0f} MB/step and {n_params * 2 / 1e6:.0f} MB of bf16 weights > 126 MB L2"},
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": 2 * B * 3 * 224 * 224 * 4, "d2h_bytes_per_step": 4,
                "api": "LearnableCompressor.fit(iterable of pinned host float32 batches): H2D of batch i + 1 under step i, every step's loss copied to pinned host memory (loss_history), logged scalars read back at the end"},
        "collective": {"grad_allreduce_bytes_per_step": (n_params * 2 if not args.frozen else n_params * 4) if world > 1 else 0,
                       "logits_tile_allgather_bytes_per_rank": (2 * B) ** 2 * 4 if (world > 1 and args.distortion == "BINCE") else 0,
                       "exposed_ms_per_step": exposed},
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (forward, dgrad and wgrad of every ViT linear)",
                     "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": None,
                     "peak_source": pk["source"], "algorithmic_gflop_per_sample": views * passes * gflop_img,
                     "reference_gflop_per_sample": views * passes * GFLOP_GEMM_PER_IMG,
                     "note": "whole-step FLOP rate per GPU (step time includes every non-GEMM kernel and the collectives)"},
        "loss": logged.get("train/feat/loss"), "rate_bits": logged.get("train/feat/rate"),
        "clocks": clocks.summary(),
    }
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
