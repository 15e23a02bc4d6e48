"""GPU numerics of the encoder's fine-tuning path (unfrozen CLIP ViT): each backward kernel against torch
autograd in fp32, and the whole hand-written backward against autograd through the fp32 CPU oracle
(oracle/vit.py) on the same weights and inputs.  Tolerances: 16-bit activations / gradients, stated below."""
import math

import pytest
import torch

from oracle import vit as ovit

pytestmark = pytest.mark.gpu
DTYPES = [torch.bfloat16, torch.float16]


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def test_transpose_pad_and_colsum(cuda_device):
    from lossyless_b200 import _native as N
    from lossyless_b200.clip_vit import _colsum, _transpose_pad

    g = torch.Generator(device=cuda_device).manual_seed(0)
    for R, C in ((150, 768), (97, 100), (64, 64), (1, 33)):
        t = torch.randn(R, C, generator=g, device=cuda_device).to(torch.bfloat16)
        out = _transpose_pad(t)
        Rp = (R + 63) // 64 * 64
        assert out.shape == (C, Rp)
        assert torch.equal(out[:, :R], t.t())
        assert not out[:, R:].any()
        s = _colsum(t)
        assert torch.allclose(s, t.float().sum(0), rtol=1e-5, atol=1e-4)
    with pytest.raises(N.NativeError):
        N.call("ll_transpose_pad", N.ptr(t), N.ptr(t), 4, 4, 2, N.stream())  # Rpad < R
    # transpose with the column sums (bias gradient) taken on the way
    for dtype in DTYPES:
        t = torch.randn(1237, 768, generator=g, device=cuda_device).to(dtype)
        out = torch.empty((768, 1280), dtype=dtype, device=cuda_device)
        db = torch.zeros(768, device=cuda_device)
        N.call("ll_transpose_pad_colsum", N.ptr(t), N.ptr(out), 1237, 768, 1280, N.dtype_code(t), N.ptr(db), N.stream())
        assert torch.equal(out[:, :1237], t.t()) and not out[:, 1237:].any()
        assert torch.allclose(db, t.float().sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("dtype", DTYPES)
def test_quickgelu(cuda_device, dtype):
    from lossyless_b200 import _native as N

    g = torch.Generator(device=cuda_device).manual_seed(1)
    u = (torch.randn(1000, 3072, generator=g, device=cuda_device) * 3).to(dtype)
    gf = torch.randn(1000, 3072, generator=g, device=cuda_device).to(dtype)
    f, gu = torch.empty_like(u), torch.empty_like(u)
    N.call("ll_quickgelu_forward", N.ptr(u), N.ptr(f), u.numel(), N.dtype_code(u), N.stream())
    N.call("ll_quickgelu_backward", N.ptr(u), N.ptr(gf), N.ptr(gu), u.numel(), N.dtype_code(u), N.stream())
    ur = u.float().requires_grad_(True)
    fr = ur * torch.sigmoid(1.702 * ur)
    fr.backward(gf.float())
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3
    assert (f.float() - fr).abs().max().item() < tol * fr.abs().max().item()
    assert (gu.float() - ur.grad).abs().max().item() < tol * ur.grad.abs().max().item()
    # a length that is not a multiple of the 8-element vector
    f2 = torch.empty(13, dtype=dtype, device=cuda_device)
    N.call("ll_quickgelu_forward", N.ptr(u), N.ptr(f2), 13, N.dtype_code(u), N.stream())
    assert torch.equal(f2, f.reshape(-1)[:13])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_backward(cuda_device, dtype, with_res):
    from lossyless_b200.clip_vit import _ln_backward

    g = torch.Generator(device=cuda_device).manual_seed(2)
    rows = 1237
    x = (torch.randn(rows, 768, generator=g, device=cuda_device) * 2 + 0.3).to(dtype)
    gh = torch.randn(rows, 768, generator=g, device=cuda_device).to(dtype)
    res = torch.randn(rows, 768, generator=g, device=cuda_device).to(dtype) if with_res else None
    w = 1 + 0.2 * torch.randn(768, generator=g, device=cuda_device)
    b = torch.randn(768, generator=g, device=cuda_device)
    gx, gw, gb = _ln_backward(x, gh, w, res)
    xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (768,), wr, br, 1e-5).backward(gh.float())
    ref = xr.grad + (res.float() if with_res else 0)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert (gx.float() - ref).abs().max().item() < tol * ref.abs().max().item()
    assert _rel(gw, wr.grad) < 1e-4 and _rel(gb, br.grad) < 1e-4  # fp32 accumulation of the 16-bit inputs


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("S,B,H", [(50, 5, 12), (64, 2, 3), (17, 3, 2)])
def test_attention_backward(cuda_device, dtype, S, B, H, impl):
    """impl 0 = ldmatrix-fed mma.sync kernel (P / dS rounded to the 16-bit type between the products), 1 = fp32
    kernel, 2 = the first mma.sync kernel (operands staged in both orientations)"""
    from lossyless_b200 import _native as N

    d = 64
    g = torch.Generator(device=cuda_device).manual_seed(S)
    qkv = torch.randn(B * S, 3 * H * d, generator=g, device=cuda_device).to(dtype)
    go = torch.randn(B * S, H * d, generator=g, device=cuda_device).to(dtype)
    gqkv = torch.full_like(qkv, float("nan"))
    N.call("ll_attention_backward_set_impl", impl)
    try:
        N.call("ll_attention_backward", N.ptr(qkv), N.ptr(go), N.ptr(gqkv), B, S, H, N.dtype_code(qkv), N.stream())
    finally:
        N.call("ll_attention_backward_set_impl", 0)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = ref_in.reshape(B, S, 3, H, d).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
    (att @ v).permute(0, 2, 1, 3).reshape(B * S, H * d).backward(go.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert torch.isfinite(gqkv.float()).all()
    assert (gqkv.float() - ref_in.grad).abs().max().item() < tol * ref_in.grad.abs().max().item()


def _trainable_vit(cuda_device, layers, dtype):
    from lossyless_b200.clip_vit import VisionTransformer

    cfg = ovit.vit_config(layers=layers)
    sd = ovit.init_state_dict(0, cfg)
    vit = VisionTransformer(224, 32, 768, layers, 12, 512)
    vit.load_state_dict(sd)
    return vit.to(cuda_device).unfreeze_(dtype), sd, cfg


@pytest.mark.parametrize("dtype", DTYPES)
def test_encoder_gradients_match_oracle_autograd(cuda_device, dtype):
    """2-block encoder, 3 images: z and every parameter gradient of sum(z * r) against autograd through
    the fp32 oracle.  Tolerance: relative Frobenius error per tensor (16-bit activations and gradients)."""
    vit, sd, cfg = _trainable_vit(cuda_device, 2, dtype)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(3, 3, 224, 224, generator=gen)
    r = torch.randn(3, 512, generator=gen)
    z = vit(x.to(cuda_device))
    assert z.dtype == torch.float32 and z.requires_grad
    (z * r.to(cuda_device)).sum().backward()

    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    z_ref = ovit.forward(ref, x, cfg)
    (z_ref * r).sum().backward()
    tol_z, tol_g = (4e-2, 6e-2) if dtype == torch.bfloat16 else (6e-3, 1.5e-2)
    assert _rel(z.detach().cpu(), z_ref.detach()) < tol_z
    worst = {}
    for name, p in vit.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, name
        worst[name] = _rel(p.grad.cpu(), ref[name].grad)
    bad = {k: v for k, v in worst.items() if not v < tol_g}
    assert not bad, bad


def test_weight_gradient_paths_agree(cuda_device, monkeypatch):
    """ll_gemm_tn (MN-major operands, default) and the transposed-copy path give the same weight gradients up to
    the 16-bit rounding of the result."""
    from lossyless_b200 import clip_vit

    grads = {}
    for impl in ("tn", "transpose"):
        monkeypatch.setattr(clip_vit, "WGRAD_IMPL", impl)
        vit, sd, cfg = _trainable_vit(cuda_device, 2, torch.bfloat16)
        gen = torch.Generator().manual_seed(5)
        x = torch.randn(3, 3, 224, 224, generator=gen).to(cuda_device)
        r = torch.randn(3, 512, generator=gen).to(cuda_device)
        (vit(x) * r).sum().backward()
        grads[impl] = {n: p.grad.clone() for n, p in vit.named_parameters()}
    for n in grads["tn"]:
        assert _rel(grads["tn"][n], grads["transpose"][n]) < 1e-2, n


def test_full_depth_gradient_direction(cuda_device):
    """ViT-B/32 (12 blocks), 2 images, bf16 compute: the flattened gradient points where autograd's does."""
    vit, sd, cfg = _trainable_vit(cuda_device, 12, torch.bfloat16)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 224, 224, generator=gen)
    r = torch.randn(2, 512, generator=gen)
    (vit(x.to(cuda_device)) * r.to(cuda_device)).sum().backward()
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    (ovit.forward(ref, x, cfg) * r).sum().backward()
    a = torch.cat([p.grad.reshape(-1).cpu() for _, p in vit.named_parameters()])
    b = torch.cat([ref[n].grad.reshape(-1) for n, _ in vit.named_parameters()])
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    assert cos > 0.995, cos
    assert 0.9 < (a.norm() / b.norm()).item() < 1.1


def test_optimizer_step_reaches_both_paths(cuda_device):
    """After an optimiser step the 16-bit compute copies (fine-tuning path) and the frozen runner's weight
    views (inference path) are rebuilt from the updated float32 master parameters."""
    vit, sd, cfg = _trainable_vit(cuda_device, 2, torch.bfloat16)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3)).to(cuda_device)
    with torch.no_grad():
        z0 = vit(x).float()
    opt = torch.optim.SGD(vit.parameters(), lr=0.05)
    z = vit(x)
    z.square().mean().backward()
    opt.step()
    with torch.no_grad():
        z1 = vit(x).float()
    z1_train = vit(x).detach()
    assert (z1 - z0).abs().max().item() > 1e-3          # the frozen runner saw the update
    assert _rel(z1, z1_train) < 5e-2                    # and both paths agree on the new weights
    assert z1_train.square().mean().item() < z0.square().mean().item()  # the step went downhill


@pytest.mark.parametrize("distortion", ["lossy_Z", "BINCE"])
def test_rd_step_with_unfrozen_encoder(cuda_device, distortion):
    """config 3 with the CLIP tower unfrozen: one LearnableCompressor step back-propagates into the
    encoder, the bottleneck and (BINCE) the projector; every gradient is finite."""
    from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams

    hp = default_clip_hparams(distortion=distortion, beta=1e-2, max_steps=100, n_data=5000, is_frozen=False)
    model = LearnableCompressor(hp).to(cuda_device).train()
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(4, 3, 224, 224, generator=gen).to(cuda_device)
    x_pos = torch.randn(4, 3, 224, 224, generator=gen).to(cuda_device)
    loss, logs, _ = model.step((x, (None, x_pos)))
    loss.backward()
    enc = dict(model.p_ZlX.named_parameters())
    assert len(enc) > 100
    if distortion == "BINCE":  # lossy_Z with is_endToEnd=False detaches the encoder from the rate; its distortion is |z_hat - z|
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.values())
        assert sum(float(p.grad.abs().sum()) for p in enc.values()) > 0
    assert model.rate_estimator.entropy_bottleneck._matrix0.grad is not None
    assert math.isfinite(float(loss))
