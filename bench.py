#!/usr/bin/env python
"""Benchmark of the encode hot path: img/sec through CLIP ViT-B/32 + EntropyBottleneck + rANS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one batch of synthetic CLIP-normalised 3x224x224 fp16 images through
``clip_compressor_b005`` (ViT-B/32 -> process_z_in -> quantise -> rANS -> container records),
i.e. the body of the reference's ``compress_dataset`` loop (hub/compressor.py:186-189).
Prints ONE JSON line on rank 0 (contract in the task statement):
  value     images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through ``ClipCompressor.compress_stream`` (compress_dataset's loop) with pinned HOST
            buffers: H2D of the images and D2H of the records inside the timed region
  roofline  the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / event-timed GEMM time,
            against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the same path on the host cores, bounded sample (N=1)
With ``--impl reference`` the CPU restatement of the reference's own call pattern is timed
instead (the reference itself cannot be installed: compressai / clip are not available).
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


# ----------------------------------------------------------------------------- CPU legs
def cpu_reference_pipeline(n_images, seed=0, faithful=True):
    """The reference's path restated on the CPU (oracle): fp32 ViT-B/32 (torch CPU, all host
    threads) -> process_z_in -> quantise -> one rANS call per image.  faithful=True mirrors
    compressai 1.1.5's call pattern (per-image Python loop, `.tolist()` of the symbols and of the
    whole CDF table, hub/compressor.py:98).  Returns (seconds, bits_per_img)."""
    import numpy as np
    import torch

    from oracle import eb as oeb
    from oracle import rans
    from oracle import vit as ovit

    st = _cpu_state()
    x = torch.randn(n_images, 3, 224, 224, generator=torch.Generator().manual_seed(seed)).half().float()
    t0 = time.perf_counter()
    with torch.no_grad():
        z = ovit.forward(st["vit"], x).half()  # reference: fp16 z
    z_in = oeb.process_z_in(z, st["sd"]["biasing"], st["sd"]["scaling"])
    sym = torch.round(z_in - oeb.medians(st["eb"])[None]).int()
    tab = st["tab"]
    nbytes = 0
    if faithful:
        idx = list(range(512))
        cdf_t, len_t, off_t = torch.from_numpy(tab["cdf"]), torch.from_numpy(tab["cdf_length"]), torch.from_numpy(tab["offset"])
        for i in range(n_images):
            s = rans.encode_with_indexes(sym[i].reshape(-1).int().tolist(), idx, np.array(cdf_t.tolist(), dtype=np.int32),
                                         len_t.reshape(-1).int().tolist(), off_t.reshape(-1).int().tolist())
            nbytes += len(s) + 4
    else:
        nbytes = sum(len(s) + 4 for s in rans.encode_batch(sym.numpy(), tab["cdf"], tab["cdf_length"], tab["offset"]))
    return time.perf_counter() - t0, 8.0 * nbytes / n_images


_CPU = {}


def _cpu_state():
    if not _CPU:
        import torch

        from oracle import eb as oeb
        from oracle import vit as ovit

        sd = torch.load(os.path.join(ROOT, "lossyless_b200", "hub", "weights", "beta5e-02_factorized_rate.pt"),
                        map_location="cpu", weights_only=True)
        eb = oeb.split_state_dict(sd)
        _CPU.update(sd=sd, eb=eb, tab=oeb.update(eb), vit=ovit.init_state_dict(0))
    return _CPU


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path (oracle port; compressai / clip cannot be
    installed) timed on this box's host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    n = args.cpu_images
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pipeline(min(n, 4))
    t, bits = 0.0, 0.0
    for s in range(args.steps):
        dt, bits = cpu_reference_pipeline(n, seed=s)
        t += dt
    value = n * args.steps / t
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n} images per step on the host CPU", "bits_per_img": bits},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {n} images: fp32 torch-CPU ViT-B/32 + per-image rANS "
                                   f"(compressai call pattern), os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (average over one block's four GEMMs) from the
    committed ncu --set full capture; None if the summary is not in the tree."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_vit_gemms_dram_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["per_launch_dram_bytes_layer_average"])
    except (OSError, KeyError, ValueError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1512, help="images per step per GPU (one 1512-image encoder pass = 4 full rounds of 74 CTA pairs x 256 token rows)")
    ap.add_argument("--chunk", type=int, default=0, help="images per internal ViT pass (0 = library default)")
    ap.add_argument("--cpu-images", type=int, default=16, help="images per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse-ln", action="store_true", help="A/B: run ln_1 / ln_2 as separate kernels")
    ap.add_argument("--attn-impl", type=int, default=-1, help="A/B: ll_attention_set_impl (0 cp.async, 1 tcgen05, 2 TMA); -1 = library default")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    warnings.filterwarnings("ignore", message="no CLIP weights given")
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the measured path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import hubconf

    comp, _ = hubconf.clip_compressor_b005(device=f"cuda:{local_rank}")
    if args.chunk:
        comp.clip.set_chunk(args.chunk)
    if args.no_fuse_ln:
        comp.clip.set_fuse_layernorm(False)
    if args.attn_impl >= 0:
        from lossyless_b200 import _native as _N

        _N.call("ll_attention_set_impl", args.attn_impl)
    B = args.batch
    # encode shards by image: every rank codes its own independent batch, no collective on the path
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    X = torch.randn(B, 3, 224, 224, generator=gen, device=dev, dtype=torch.float32).half()
    X_host = X.cpu().pin_memory()

    def step_device():
        return comp._compress_records(X, mode=1)

    for _ in range(args.warmup):
        packed, offsets, lens = step_device()
    barrier()
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
    def e2e_run(x_host, steps):
        n = 0
        for rec, b in comp.compress_stream(x_host for _ in range(steps)):
            n += b
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

    # ---- roofline leg: same steps with every GEMM launch bracketed by events
    comp.clip.profile_begin()
    for _ in range(max(2, min(args.steps, 5))):
        step_device()
    gemm_ms, gemm_flops, gemm_launches = comp.clip.profile_end()
    breakdown = {k: v / max(2, min(args.steps, 5)) for k, v in comp.clip.profile_breakdown().items()}
    pk = peaks()
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    vit_launches = comp.clip.launches(B)
    line = {
        "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 (fp32 accumulate; integer rANS)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "beta": 0.05, "vit_weights": "seeded random init (no pretrained CLIP offline)",
                   "l2": f"inputs {B * 3 * 224 * 224 * 2 / 1e6:.0f} MB/step > 126 MB L2, no flush needed",
                   "bits_per_img": bits_per_img, "sharding": "independent images per rank, no collective"},
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": B * 3 * 224 * 224 * 2, "d2h_bytes_per_step": total_bytes + 8,
                "input": "pinned host fp16 [B,3,224,224] CLIP-normalised tensors (what the reference's loop body receives)",
                "api": "ClipCompressor.compress_stream (compress_dataset's loop)", "h2d_link_gbps": h2d_gbps, "blocking_per_batch_call": e2e_blocking},
        "e2e_raw_uint8": {"value": e2e_raw, "unit": "img/s", "h2d_bytes_per_step": B * 96 * 96 * 3,
                          "input": "pinned host uint8 [B,96,96,3] STL10-shaped images; resize/crop/normalise on the GPU"},
        "gpu_launches": args.steps * (vit_launches + 1 + 2),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all ViT linears)", "achieved": achieved,
                     "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": ncu_traffic(),
                     "traffic_unit": "bytes of DRAM read + write per GEMM launch (ncu --set full capture, profiles/r01_ncu_vit_gemms_dram_traffic.json)",
                     "peak_source": pk["source"], "gemm_ms_per_step": gemm_ms / max(2, min(args.steps, 5)),
                     "gemm_launches_per_step": gemm_launches // max(2, min(args.steps, 5)),
                     "algorithmic_gflop_per_img": GFLOP_GEMM_PER_IMG,
                     "kernel_ms_per_step": breakdown},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        n = args.cpu_images
        cpu_reference_pipeline(4)
        t, _ = cpu_reference_pipeline(n)
        reps = max(1, min(8, int(12.0 / max(t, 1e-3))))
        tt = t + sum(cpu_reference_pipeline(n, seed=i + 1)[0] for i in range(reps - 1))
        line["cpu_baseline"] = {"value": n * reps / tt, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{reps} x {n} images: fp32 torch-CPU ViT-B/32 + per-image rANS oracle, os.cpu_count()={os.cpu_count()}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
