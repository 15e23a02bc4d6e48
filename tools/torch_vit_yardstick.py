"""Library yard-stick: CLIP ViT-B/32 forward written in plain torch -- F.linear (cuBLASLt), F.layer_norm,
F.scaled_dot_product_attention (flash / mem-efficient kernels), fp16 weights and activations, LayerNorm in fp32 as
CLIP's own LayerNorm does -- on the same box, same batch, same weights and images as bench.py, next to this
package's hand-written encoder.  Reports img/s of both and, against a float32 run of the same torch model, the
|dz| and the symbol-flip rate each 16-bit encoder induces at the three published bottlenecks.

    python tools/torch_vit_yardstick.py [--batch 1512] [--steps 20] [--out profiles/r02_torch_vit_yardstick.txt]
"""
import argparse
import os
import sys
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def torch_vit(sd, x, dtype, heads=12):
    """sd: CLIP visual-tower state dict already in `dtype`; x [B,3,224,224] in `dtype`."""
    W = sd["class_embedding"].shape[0]
    B = x.shape[0]

    def ln(t, p):  # CLIP LayerNorm: compute in fp32, cast back
        return F.layer_norm(t.float(), (W,), sd[p + ".weight"].float(), sd[p + ".bias"].float(), 1e-5).to(dtype)

    t = F.conv2d(x, sd["conv1.weight"], stride=32).reshape(B, W, -1).permute(0, 2, 1)
    t = torch.cat([sd["class_embedding"].expand(B, 1, W), t], dim=1) + sd["positional_embedding"]
    t = ln(t, "ln_pre")
    S, d = t.shape[1], W // heads
    i = 0
    while f"transformer.resblocks.{i}.ln_1.weight" in sd:
        p = f"transformer.resblocks.{i}."
        qkv = F.linear(ln(t, p + "ln_1"), sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = (u.reshape(B, S, heads, d).transpose(1, 2) for u in qkv.split(W, dim=-1))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, W)
        t = t + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = F.linear(ln(t, p + "ln_2"), sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)
        t = t + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        i += 1
    return ln(t[:, 0, :], "ln_post") @ sd["proj"]


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import bench
    import hubconf

    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    comp, _ = hubconf.clip_compressor_b005(device="cuda:0")
    sd16 = {k: v.detach() for k, v in comp.clip.state_dict().items()}
    sd32 = {k: v.float() for k, v in sd16.items()}
    X = bench.bench_images(args.batch).to(dev)
    lines = [f"# ViT-B/32 encoder forward, batch {args.batch}, {torch.cuda.get_device_name(0)}, torch {torch.__version__}; same fp16 weights and images"]
    with torch.no_grad():
        ms_ours = timed(lambda: comp.clip(X), args.steps)
        ms_torch = timed(lambda: torch_vit(sd16, X, torch.float16), args.steps)
        lines.append(f"lossyless_b200 (tcgen05 GEMMs, fused epilogues, hand-written attention): {ms_ours:8.3f} ms  {args.batch / ms_ours * 1e3:9.0f} img/s")
        lines.append(f"torch fp16 (cuBLASLt F.linear + SDPA + ATen LayerNorm/QuickGELU)        : {ms_torch:8.3f} ms  {args.batch / ms_torch * 1e3:9.0f} img/s")
        for b in (128, 32):
            xs = X[:b]
            lines.append(f"batch {b:4d}: ours {timed(lambda: comp.clip(xs), 50):7.3f} ms   torch fp16 {timed(lambda: torch_vit(sd16, xs, torch.float16), 50):7.3f} ms")
        n = min(256, args.batch)
        z32 = torch.cat([torch_vit(sd32, X[i:i + 64].float(), torch.float32) for i in range(0, n, 64)])
        z_ours, z_t16 = comp.clip(X[:n]).float(), torch_vit(sd16, X[:n], torch.float16).float()
        lines.append(f"|z - z_fp32| on {n} images (z rms {z32.pow(2).mean().sqrt():.3f}): ours max {float((z_ours - z32).abs().max()):.2e} mean {float((z_ours - z32).abs().mean()):.2e}"
                     f"   torch fp16 max {float((z_t16 - z32).abs().max()):.2e} mean {float((z_t16 - z32).abs().mean()):.2e}")
        for tag in ("beta1e-01", "beta5e-02", "beta1e-02"):
            sd = torch.load(os.path.join(ROOT, "lossyless_b200", "hub", "weights", f"{tag}_factorized_rate.pt"), map_location=dev, weights_only=True)
            med = sd["entropy_bottleneck.quantiles"][:, 0, 1]

            def sym(z):
                return torch.round((z.half().float() + sd["biasing"]) * sd["scaling"].exp() - med)

            s32 = sym(z32)
            lines.append(f"{tag}: symbol flips vs fp32  ours {float((sym(z_ours) != s32).float().mean()):.4f}   torch fp16 {float((sym(z_t16) != s32).float().mean()):.4f}"
                         f"   (fp16 rounding of an exact z alone: {float((torch.round((z32 + sd['biasing']) * sd['scaling'].exp() - med) != s32).float().mean()):.4f})")
    print("\n".join(lines))
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
