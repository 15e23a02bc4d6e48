"""SURVEY.md 8(d) configuration 3: the rate-distortion training step (CLIP ViT-B/32 encoder frozen as in
config/finetune/freezer.yaml, or unfrozen; rate = H_factorized, distortion = lossy_Z or BINCE), x, x_pos ~ N(0,1)
device-resident, per-GPU batch in {64, 256}, AdamW 1e-3 / 3e-4, grad-clip 3.  One process per GPU:

    python tools/train_bench.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/train_bench.py      # data parallel: gradient all-reduce + logits-tile all-gather

Prints one line per (distortion, batch): ms/step (CUDA events, max over ranks) and images/s over all ranks
(BINCE encodes two views per sample; images/s counts samples).
"""
import argparse
import os
import sys
import warnings

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", choices=["", "frozen", "unfrozen"])
    ap.add_argument("--batches", default="64,256", help="per-GPU batch sizes")
    ap.add_argument("--wgrad", default="", choices=["", "tn", "transpose"], help="A/B: weight-gradient GEMM path (clip_vit.WGRAD_IMPL)")
    args = ap.parse_args()
    if args.wgrad:
        from lossyless_b200 import clip_vit

        clip_vit.WGRAD_IMPL = args.wgrad
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lines = [f"# RD training step, ViT-B/32 encoder frozen (config/finetune/freezer.yaml) or unfrozen (bf16 compute, fp32 master, bf16 gradient buckets), "
             f"{world} x {torch.cuda.get_device_name(local)}; {args.steps} timed steps after {args.warmup} warm-ups"]
    sizes = [int(b) for b in args.batches.split(",")]
    cases = [(d, B, True) for d in ("lossy_Z", "BINCE") for B in sizes] + [("BINCE", B, False) for B in sizes]
    for distortion, B, frozen in cases:
        if args.only and args.only != ("frozen" if frozen else "unfrozen"):
            continue
        for _ in (0,):
            torch.manual_seed(1234 + rank)
            hp = default_clip_hparams(distortion=distortion, beta=1e-2, max_steps=1000, n_data=50000, is_frozen=frozen)
            model = LearnableCompressor(hp).to(dev).train()
            fit_kw = {} if frozen else dict(grad_comm_dtype=torch.bfloat16)
            x = torch.randn(B, 3, 224, 224, device=dev)
            x_pos = torch.randn(B, 3, 224, 224, device=dev)
            batch = (x, (None, x_pos))
            # LearnableCompressor.fit = the reference's Trainer.fit loop for this path: per batch one pass per
            # optimiser (main, coder), gradient all-reduce when torch.distributed is up, clip 3, step
            model.fit([batch] * args.warmup, **fit_kw)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            logged = model.fit([batch] * args.steps, **fit_kw)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                # unfrozen BINCE: 2 views x (forward + backward) = 6 x the forward FLOPs of one image per sample
                tf = "" if frozen else f"  {world * B * 6 * 8.7255e9 / ms.item() / 1e9:7.1f} TFLOP/s (6 x 8.73 GFLOP per sample)"
                lines.append(f"{distortion:8s} {'frozen  ' if frozen else 'unfrozen'} per-GPU batch {B:4d}: {ms.item():8.3f} ms/step  {world * B / ms.item() * 1e3:9.0f} img/s{tf}")
                print(lines[-1], flush=True)
            del model
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
