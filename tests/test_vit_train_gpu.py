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


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_embed_tokens_and_adjoint(cuda_device, dtype):
    """ll_embed_tokens == cat([class_embedding, patch], 1) + positional_embedding in the 16-bit type (bit-equal: one
    rounding per element either way); ll_embed_backward returns the patch rows of the gradient bit-equal and the
    embeddings' gradients (batch sums, float32) within summation-order error, accumulating into its outputs."""
    from lossyless_b200 import _native as N

    gen = torch.Generator().manual_seed(3)
    for B, S, W in ((37, 50, 768), (3, 5, 64), (300, 50, 768)):
        patch = torch.randn(B, S - 1, W, generator=gen).to(cuda_device, dtype)
        cls = torch.randn(W, generator=gen).to(cuda_device, dtype)
        pos = torch.randn(S, W, generator=gen).to(cuda_device, dtype)
        tok = torch.empty(B, S, W, device=cuda_device, dtype=dtype)
        N.call("ll_embed_tokens", N.ptr(patch), N.ptr(cls), N.ptr(pos), N.ptr(tok), B, S, W, N.dtype_code(tok), N.stream())
        assert torch.equal(tok, torch.cat([cls.expand(B, 1, W), patch], 1) + pos)
        g_tok = torch.randn(B, S, W, generator=gen).to(cuda_device, dtype)
        g_patch = torch.empty(B, S - 1, W, device=cuda_device, dtype=dtype)
        g_pos = torch.ones(S, W, device=cuda_device)
        g_cls = torch.full((W,), 2.0, device=cuda_device)
        N.call("ll_embed_backward", N.ptr(g_tok), N.ptr(g_patch), N.ptr(g_pos), N.ptr(g_cls), B, S, W, N.dtype_code(g_tok), N.stream())
        assert torch.equal(g_patch, g_tok[:, 1:])
        assert torch.allclose(g_pos, 1.0 + g_tok.float().sum(0), rtol=1e-5, atol=1e-4)
        assert torch.allclose(g_cls, 2.0 + g_tok[:, 0].float().sum(0), rtol=1e-5, atol=1e-4)


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
        monkeypatch.setattr(clip_vit, "NATIVE_BLOCKS", impl == "tn")  # the native block runner always takes ll_gemm_tn
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


# ---------------------------------------------------------------------------------------------- flat training state
@pytest.mark.parametrize("g_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_flat_optimizer_kernels_vs_torch(cuda_device, g_dtype):
    """ll_sumsq / ll_adamw_flat / ll_cast_flat against torch: |g|^2, three AdamW steps (torch.optim.AdamW semantics,
    gradient scale read from device memory) and the 16-bit compute copy the kernel writes alongside."""
    from lossyless_b200 import _native as N

    n = 8 * 12345
    gen = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=gen)
    p = p0.clone().to(cuda_device)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    w16 = torch.empty(n, dtype=torch.bfloat16, device=cuda_device)
    ref = torch.nn.Parameter(p0.clone().to(cuda_device))
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=3e-2, betas=(0.9, 0.999), eps=1e-8)
    scratch = torch.empty(int(N.load().ll_sumsq_scratch_floats()), dtype=torch.float32, device=cuda_device)
    scale = torch.tensor([0.37], device=cuda_device)
    for step in range(1, 4):
        g = (torch.randn(n, generator=gen) * 0.1).to(g_dtype).to(cuda_device)
        out = torch.full((1,), 5.0, device=cuda_device)
        N.call("ll_sumsq", N.ptr(g), n, N.dtype_code(g), N.ptr(scratch), N.ptr(out), 0, N.stream())
        assert float(out) == pytest.approx(float(g.double().square().sum()), rel=1e-5)
        N.call("ll_sumsq", N.ptr(g), n, N.dtype_code(g), N.ptr(scratch), N.ptr(out), 1, N.stream())
        assert float(out) == pytest.approx(2 * float(g.double().square().sum()), rel=1e-5)
        N.call("ll_adamw_flat", N.ptr(p), N.ptr(g), N.dtype_code(g), N.ptr(m), N.ptr(v), N.ptr(w16), N.LL_BF16, n,
               1e-3, 0.9, 0.999, 1e-8, 3e-2, step, N.ptr(scale), N.stream())
        ref.grad = g.float() * 0.37
        opt.step()
        assert torch.allclose(p, ref.detach(), rtol=2e-6, atol=2e-7), (p - ref.detach()).abs().max()
        assert torch.equal(w16, p.bfloat16())
    st = opt.state[ref]
    assert torch.allclose(m, st["exp_avg"], rtol=1e-5, atol=1e-8) and torch.allclose(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-10)
    h = torch.empty(n, dtype=torch.float16, device=cuda_device)
    N.call("ll_cast_flat", N.ptr(p), N.ptr(h), N.LL_F16, n, N.stream())
    assert torch.equal(h, p.half())


def test_flat_mode_gradients_equal_autograd_mode(cuda_device):
    """Flat training state: the backward pass writes every gradient into the flat buffers (16-bit for the GEMM-produced
    ones); they are the values the autograd-returning mode yields, the parameters alias the flat master, and the
    16-bit compute views alias the flat 16-bit buffer."""
    vit, sd, cfg = _trainable_vit(cuda_device, 2, torch.bfloat16)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 224, 224, generator=gen).to(cuda_device)
    r = torch.randn(3, 512, generator=gen).to(cuda_device)
    z_ref = vit(x)
    (z_ref * r).sum().backward()
    want = {n: p.grad.clone() for n, p in vit.named_parameters()}
    vit.zero_grad(set_to_none=True)
    vit.enable_flat_training_(True)
    z = vit(x)
    assert torch.equal(z, z_ref)
    (z * r).sum().backward()
    flat = vit.flat_state()
    assert all(p.grad is None for p in vit.parameters())
    for n, p in vit.named_parameters():
        got = flat.grad_view(n).float()
        assert got.shape == p.shape
        if flat.slots[n][0] == "mat":
            assert torch.equal(got, want[n]), n          # same GEMM, same 16-bit result
        else:
            assert _rel(got, want[n]) < 1e-4, n         # float32 atomics: order may differ
        assert p.data_ptr() == flat.view(flat.master, n).data_ptr()
    # a second backward overwrites (does not accumulate)
    (vit(x) * r).sum().backward()
    assert _rel(flat.grad_view("ln_pre.weight"), want["ln_pre.weight"]) < 1e-4
    vit.enable_flat_training_(False)
    (vit(x) * r).sum().backward()
    assert _rel(vit.ln_pre.weight.grad, want["ln_pre.weight"]) < 1e-4


def test_flat_adamw_matches_torch_adamw(cuda_device):
    """helpers.FlatAdamW (native flat-buffer AdamW + device-side global-norm clip for the tower, torch's fused AdamW for
    the other parameters) against torch.optim.AdamW + torch clip_grad_norm_ on the same model, three steps."""
    from lossyless_b200.helpers import FlatAdamW, get_optimizer

    gen = torch.Generator().manual_seed(7)
    x = torch.randn(4, 3, 224, 224, generator=gen).to(cuda_device)
    r = torch.randn(4, 512, generator=gen).to(cuda_device)
    results = {}
    for kind in ("torch", "flat"):
        vit, sd, cfg = _trainable_vit(cuda_device, 2, torch.bfloat16)
        head = torch.nn.Linear(512, 512).to(cuda_device)
        with torch.no_grad():
            head.weight.copy_(torch.eye(512, device=cuda_device) * 0.5), head.bias.zero_()
        params = list(vit.parameters()) + list(head.parameters())
        if kind == "torch":
            opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=3e-8)
        else:
            opt = get_optimizer(params, "AdamW", lr=1e-3, weight_decay=3e-8)
            assert isinstance(opt, FlatAdamW) and vit._flat_enabled
        norms, first = [], None
        for it in range(3):
            opt.zero_grad(set_to_none=True)
            if kind == "flat":
                opt.prepare_backward(sync=False)
            (head(vit(x)) * r).sum().backward()
            if kind == "torch":
                norms.append(float(torch.nn.utils.clip_grad_norm_(params, 3.0)))
            else:
                opt.finish_backward()
                norms.append(float(opt.clip_grad_norm_(3.0)))
            opt.step()
            if it == 0:
                first = {n: p.detach().clone() for n, p in list(vit.named_parameters()) + [("head.weight", head.weight)]}
        with torch.no_grad():
            results[kind] = (norms, {n: p.detach().clone() for n, p in vit.named_parameters()}, head.weight.detach().clone(), vit(x).float(), first)
    (na, pa, ha, za, p1a), (nb, pb, hb, zb, p1b) = results["torch"], results["flat"]
    assert na[0] == pytest.approx(nb[0], rel=1e-4) and na[0] > 3.0  # same global norm, and the clip was active
    # after ONE step (identical gradients in both runs) the parameters agree to float32 rounding
    for n in p1a:
        assert _rel(p1a[n], p1b[n]) < 2e-6, n
    # later steps see 16-bit weights that may differ in the last bit, and Adam turns a near-zero gradient entry's noise
    # into a full +-lr step: compare against the distance travelled (3 steps x lr = 3e-3)
    assert na == pytest.approx(nb, rel=5e-3)
    for n in pa:
        assert (pa[n] - pb[n]).abs().max().item() < 6.5e-3 and (pa[n] - pb[n]).abs().mean().item() < 1e-3, n
    assert (ha - hb).abs().max().item() < 6.5e-3 and (ha - hb).abs().mean().item() < 1e-3
    assert _rel(za, zb) < 3e-2


def test_fit_with_unfrozen_tower_uses_flat_optimizer(cuda_device):
    """LearnableCompressor.fit on config 3 (tower unfrozen): the main optimiser is a FlatAdamW, two fit() calls keep
    its state, the loss is finite and the tower's parameters move."""
    from lossyless_b200.helpers import FlatAdamW
    from lossyless_b200.learnable_compressors import LearnableCompressor, default_clip_hparams

    hp = default_clip_hparams(distortion="BINCE", beta=1e-2, max_steps=100, n_data=5000, is_frozen=False)
    torch.manual_seed(0)
    model = LearnableCompressor(hp)
    enc = model.p_ZlX.mapper.encoder
    del enc.transformer.resblocks[2:]
    enc.layers = 2
    model = model.to(cuda_device).train()
    w0 = enc.transformer.resblocks[1].mlp.c_fc.weight.detach().clone()
    gen = torch.Generator().manual_seed(1)
    batch = (torch.randn(4, 3, 224, 224, generator=gen).to(cuda_device), (None, torch.randn(4, 3, 224, 224, generator=gen).to(cuda_device)))
    logged = model.fit([batch, batch])
    assert isinstance(model._trainer_state["optimizers"][0], FlatAdamW)
    assert math.isfinite(logged["train/feat/loss"])
    flat = enc.flat_state()
    assert flat.step_count == 2
    model.fit([batch])
    assert flat.step_count == 3 and enc.flat_state() is flat
    assert (enc.transformer.resblocks[1].mlp.c_fc.weight.detach() - w0).abs().max().item() > 0
    with torch.no_grad():  # the frozen runner sees the trained weights
        z = enc(batch[0])
    assert torch.isfinite(z.float()).all()


def test_native_block_runner_equals_python_sequencing(cuda_device, monkeypatch):
    """ll_vit_block_forward / ll_vit_block_backward (one native call per block and direction) against the same kernels
    sequenced from Python: identical output, identical matrix gradients, vector gradients equal up to atomics order."""
    from lossyless_b200 import _native as N
    from lossyless_b200 import clip_vit

    out = {}
    old_mask = N.load().ll_vit_block_set_fusion(0)  # the Python sequencing has no fused epilogues: compare like with like
    monkeypatch.setattr(clip_vit, "CLS_LAST_BLOCK", False)  # ... and computes every row of the last block
    try:
        for native in (True, False):
            monkeypatch.setattr(clip_vit, "NATIVE_BLOCKS", native)
            vit, sd, cfg = _trainable_vit(cuda_device, 3, torch.bfloat16)
            gen = torch.Generator().manual_seed(11)
            x = torch.randn(5, 3, 224, 224, generator=gen).to(cuda_device)
            r = torch.randn(5, 512, generator=gen).to(cuda_device)
            z = vit(x)
            (z * r).sum().backward()
            out[native] = (z.detach().clone(), {n: p.grad.clone() for n, p in vit.named_parameters()})
    finally:
        N.load().ll_vit_block_set_fusion(old_mask)
    assert torch.equal(out[True][0], out[False][0])
    for n, g in out[True][1].items():
        if clip_vit._is_matrix(n):
            assert torch.equal(g, out[False][1][n]), n
        else:
            assert _rel(g, out[False][1][n]) < 1e-4, n


@pytest.mark.parametrize("dtype", DTYPES)
def test_block_runner_fusions_match_unfused(cuda_device, dtype):
    """The native block runner with its fused epilogues (QuickGELU backward inside the c_proj data-gradient GEMM, ...)
    against the same runner with every fusion off: same output, gradients equal up to the 16-bit rounding of the
    intermediates that no longer exist."""
    from lossyless_b200 import _native as N

    out = {}
    for mask in (0x7FFFFFFF, 0):
        old_mask = N.load().ll_vit_block_set_fusion(mask)
        try:
            vit, sd, cfg = _trainable_vit(cuda_device, 3, dtype)
            gen = torch.Generator().manual_seed(12)
            x = torch.randn(6, 3, 224, 224, generator=gen).to(cuda_device)
            r = torch.randn(6, 512, generator=gen).to(cuda_device)
            z = vit(x)
            (z * r).sum().backward()
            out[mask] = (z.detach().clone(), {n: p.grad.float().clone() for n, p in vit.named_parameters()})
        finally:
            N.load().ll_vit_block_set_fusion(old_mask)
    assert (out[0][0].float() - out[0x7FFFFFFF][0].float()).abs().max().item() < (3e-2 if dtype == torch.bfloat16 else 4e-3)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    for n, g in out[0][1].items():
        assert _rel(out[0x7FFFFFFF][1][n], g) < tol, n


@pytest.mark.parametrize("dtype", DTYPES)
def test_cls_only_last_block_training_matches_all_rows(cuda_device, dtype, monkeypatch):
    """The last block's out-projection / ln_2 / MLP -- forward, data and weight gradients -- on the class-token rows only
    (default) against the same runner computing every row: identical z (same per-element accumulation order), every
    parameter gradient equal up to the summation order of the weight-gradient reductions (B terms instead of B * S
    with zeros in between)."""
    from lossyless_b200 import clip_vit

    out = {}
    for cls in (True, False):
        monkeypatch.setattr(clip_vit, "CLS_LAST_BLOCK", cls)
        vit, sd, cfg = _trainable_vit(cuda_device, 3, dtype)
        gen = torch.Generator().manual_seed(13)
        x = torch.randn(7, 3, 224, 224, generator=gen).to(cuda_device)
        r = torch.randn(7, 512, generator=gen).to(cuda_device)
        z = vit(x)
        (z * r).sum().backward()
        out[cls] = (z.detach().clone(), {n: p.grad.float().clone() for n, p in vit.named_parameters()})
    assert torch.equal(out[True][0], out[False][0])
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3
    for n, g in out[False][1].items():
        assert _rel(out[True][1][n], g) < tol, n


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.float16])
def test_forward_views_equals_forward_of_the_concatenation(cuda_device, in_dtype):
    """VisionTransformer.forward_views([x, x_pos]) (ll_patchify per view, straight into one im2col buffer) is
    forward(torch.cat([x, x_pos])) -- values and gradients -- on the fine-tuning path and on the frozen one."""
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(3, 3, 224, 224, generator=gen).to(cuda_device).to(in_dtype)
    xp = torch.randn(5, 3, 224, 224, generator=gen).to(cuda_device).to(in_dtype)
    r = torch.randn(8, 512, generator=gen).to(cuda_device)
    out = {}
    for how in ("views", "cat"):
        vit, sd, cfg = _trainable_vit(cuda_device, 2, torch.bfloat16)
        z = vit.forward_views([x, xp]) if how == "views" else vit(torch.cat([x, xp]))
        (z * r).sum().backward()
        out[how] = (z.detach().clone(), {n: p.grad.float().clone() for n, p in vit.named_parameters()})
    assert torch.equal(out["views"][0], out["cat"][0])
    for n, g in out["cat"][1].items():
        assert torch.equal(out["views"][1][n], g) or _rel(out["views"][1][n], g) < 1e-5, n
    with torch.no_grad():  # frozen path: a plain concatenation
        assert torch.equal(vit.forward_views([x, xp]), vit(torch.cat([x, xp])))
