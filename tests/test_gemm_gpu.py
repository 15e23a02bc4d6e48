"""GPU numerics: the tcgen05 GEMM (+ fused epilogues) against a plain fp32 torch reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, W, bias=None, residual=None, act=0):
    from lossyless_b200 import _native as N

    M, K = A.shape
    D = torch.empty((M, W.shape[0]), dtype=A.dtype, device=A.device)
    N.call("ll_gemm", N.ptr(A), N.ptr(W), N.ptr(bias), N.ptr(residual), N.ptr(D), M, W.shape[0], K,
           N.dtype_code(A), act, N.stream())
    return D


def _ref(A, W, bias, residual, act):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    if residual is not None:
        y = y + residual.float()
    return y


SHAPES = [  # (M, N, K) -- the encoder's GEMMs at small and ragged M, plus BN=256 sized problems
    (1, 512, 768), (50, 768, 768), (77, 2304, 768), (128, 3072, 768), (333, 768, 3072),
    (6272, 768, 3072), (6400, 2304, 768), (12800, 768, 768), (40000, 768, 768), (40000, 3072, 768),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain(cuda_device, dtype, M, N, K):
    g = torch.Generator(device=cuda_device).manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g, device=cuda_device) * 0.5).to(dtype)
    W = (torch.randn(N, K, generator=g, device=cuda_device) * K ** -0.5).to(dtype)
    D = _gemm(A, W)
    ref = _ref(A, W, None, None, 0)
    tol = 4e-3 if dtype == torch.float16 else 2e-2
    err = (D.float() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("act,use_res", [(0, True), (1, False), (1, True), (0, False)])
def test_gemm_epilogues(cuda_device, act, use_res):
    from lossyless_b200 import _native as nat

    g = torch.Generator(device=cuda_device).manual_seed(5)
    M, Nn, K = 6400 + 17, 768, 768
    A = (torch.randn(M, K, generator=g, device=cuda_device) * 0.5).half()
    W = (torch.randn(Nn, K, generator=g, device=cuda_device) * K ** -0.5).half()
    bias = torch.randn(Nn, generator=g, device=cuda_device)
    res = torch.randn(M, Nn, generator=g, device=cuda_device).half() if use_res else None
    ref = _ref(A, W, bias, res, act)
    D = _gemm(A, W, bias, res, act)
    assert (D.float() - ref).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())
    if use_res:  # in-place residual (D aliases the residual), as the encoder uses it
        x = res.clone()
        nat.call("ll_gemm", nat.ptr(A), nat.ptr(W), nat.ptr(bias), nat.ptr(x), nat.ptr(x), M, Nn, K, 1, act, nat.stream())
        assert torch.equal(x, D)


def test_gemm_repeatable_and_back_to_back(cuda_device):
    """persistent kernel state (barrier phases, TMEM alloc/dealloc) survives many launches and
    gives bit-identical results."""
    g = torch.Generator(device=cuda_device).manual_seed(9)
    A = torch.randn(12800, 768, generator=g, device=cuda_device).half()
    W = (torch.randn(3072, 768, generator=g, device=cuda_device) * 0.03).half()
    first = _gemm(A, W)
    for _ in range(20):
        assert torch.equal(_gemm(A, W), first)


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
@pytest.mark.parametrize("use_res", [False, True])
def test_gemm_every_tile_width(cuda_device, bn, use_res):
    """each tile-N instantiation (odd epilogue group counts included) with and without the TMA-fed
    residual, several tiles per CTA so the staging buffers / barrier phases wrap."""
    from lossyless_b200 import _native as nat

    g = torch.Generator(device=cuda_device).manual_seed(bn)
    M, Nn, K = 128 * 148 * 2 + 77, 768, 768
    A = (torch.randn(M, K, generator=g, device=cuda_device) * 0.5).half()
    W = (torch.randn(Nn, K, generator=g, device=cuda_device) * K ** -0.5).half()
    bias = torch.randn(Nn, generator=g, device=cuda_device)
    res = torch.randn(M, Nn, generator=g, device=cuda_device).half() if use_res else None
    ref = _ref(A, W, bias, res, 0)
    nat.call("ll_gemm_set_tile_n", bn)
    try:
        D = _gemm(A, W, bias, res, 0)
    finally:
        nat.call("ll_gemm_set_tile_n", 0)
    assert (D.float() - ref).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("bn", [128, 192, 256])
@pytest.mark.parametrize("M,use_res,act", [(128 * 148 * 2 + 77, False, 0), (12800, True, 0), (300, False, 1), (25600, True, 1)])
def test_gemm_cta_pair(cuda_device, bn, M, use_res, act):
    """tcgen05 cta_group::2 path (256-row tiles on CTA pairs), forced on: ragged M (the peer CTA of the
    last pair partly or fully out of range), residual, QuickGELU, several tiles per pair."""
    from lossyless_b200 import _native as nat

    g = torch.Generator(device=cuda_device).manual_seed(bn + M)
    Nn, K = 768, 768
    A = (torch.randn(M, K, generator=g, device=cuda_device) * 0.5).half()
    W = (torch.randn(Nn, K, generator=g, device=cuda_device) * K ** -0.5).half()
    bias = torch.randn(Nn, generator=g, device=cuda_device)
    res = torch.randn(M, Nn, generator=g, device=cuda_device).half() if use_res else None
    ref = _ref(A, W, bias, res, act)
    nat.call("ll_gemm_set_tile_n", bn)
    nat.call("ll_gemm_set_cta_pair", 1)
    try:
        D = _gemm(A, W, bias, res, act)
        D2 = _gemm(A, W, bias, res, act)
    finally:
        nat.call("ll_gemm_set_tile_n", 0)
        nat.call("ll_gemm_set_cta_pair", -1)
    assert torch.equal(D, D2)
    assert (D.float() - ref).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())


def test_gemm_cta_pair_long_k(cuda_device):
    from lossyless_b200 import _native as nat

    g = torch.Generator(device=cuda_device).manual_seed(77)
    M, Nn, K = 20000, 768, 3072
    A = (torch.randn(M, K, generator=g, device=cuda_device) * 0.5).half()
    W = (torch.randn(Nn, K, generator=g, device=cuda_device) * K ** -0.5).half()
    nat.call("ll_gemm_set_cta_pair", 1)
    try:
        D = _gemm(A, W)
    finally:
        nat.call("ll_gemm_set_cta_pair", -1)
    ref = _ref(A, W, None, None, 0)
    assert (D.float() - ref).abs().max().item() < 6e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,Nn,act", [(50, 2304, 0), (6400 + 17, 3072, 1), (18944, 768, 0)])
def test_gemm_fused_layernorm(cuda_device, dtype, M, Nn, act):
    """ll_gemm_ln: LN(A) W^T + b from the raw rows, per-row statistics partials and host-folded weights, against
    torch's layer_norm + matmul in fp32; and the row statistics it leaves for the next LayerNorm."""
    from lossyless_b200 import _native as nat

    g = torch.Generator(device=cuda_device).manual_seed(M + Nn)
    K = 768
    A = (torch.randn(M, K, generator=g, device=cuda_device) * 0.7 + torch.randn(M, 1, generator=g, device=cuda_device)).to(dtype)
    A[:, 5] *= 30  # an outlier channel like the residual stream's
    W = (torch.randn(Nn, K, generator=g, device=cuda_device) * K ** -0.5).to(dtype)
    bias = torch.randn(Nn, generator=g, device=cuda_device)
    gamma = 1 + 0.2 * torch.randn(K, generator=g, device=cuda_device)
    beta = 0.2 * torch.randn(K, generator=g, device=cuda_device)
    ref = _ref(torch.nn.functional.layer_norm(A.float(), (K,), gamma, beta, 1e-5), W, bias, None, act)
    # the fold ll_vit_create does on the device, here in torch
    Wg = W.float() * gamma
    Wf = (Wg - Wg.mean(1, keepdim=True)).to(dtype)
    bf = bias + W.float() @ beta
    # statistics of A in three uneven parts, [parts, M] float2
    cuts = [0, 100, 101, K]
    parts = torch.stack([torch.stack([A.float()[:, a:b].sum(1), (A.float()[:, a:b] ** 2).sum(1)], -1) for a, b in zip(cuts, cuts[1:])]).contiguous()
    D = torch.empty((M, Nn), dtype=dtype, device=cuda_device)
    st_out = torch.full((Nn // 64, M, 2), float("nan"), device=cuda_device)
    nat.call("ll_gemm_ln", nat.ptr(A), nat.ptr(Wf), nat.ptr(bf), None, nat.ptr(D), M, Nn, K, nat.dtype_code(A), act,
             nat.ptr(parts), 3, nat.ptr(st_out), K, nat.stream())
    tol = 6e-3 if dtype == torch.float16 else 3e-2
    err = (D.float() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err
    # row statistics for the next LayerNorm (taken before the 16-bit rounding): mean and variance of the stored rows
    got = st_out.sum(0)
    mean, var = got[:, 0] / Nn, got[:, 1] / Nn - (got[:, 0] / Nn) ** 2
    eps = 1e-3 if dtype == torch.float16 else 8e-3
    assert (mean - D.float().mean(1)).abs().max().item() < eps
    assert ((var - D.float().var(1, unbiased=False)).abs() / D.float().var(1, unbiased=False)).max().item() < eps
    # reproducible: a second launch gives the same bits (no atomics in the statistics)
    D2, st2 = torch.empty_like(D), torch.empty_like(st_out)
    nat.call("ll_gemm_ln", nat.ptr(A), nat.ptr(Wf), nat.ptr(bf), None, nat.ptr(D2), M, Nn, K, nat.dtype_code(A), act,
             nat.ptr(parts), 3, nat.ptr(st2), K, nat.stream())
    assert torch.equal(D, D2) and torch.equal(st_out, st2)


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("Kt,Mo,No", [(2500, 768, 3072), (12800, 768, 768), (137, 2304, 768), (64, 256, 512), (9, 768, 512)])
def test_gemm_tn_weight_gradient_shape(cuda_device, dtype, Kt, Mo, No, pair):
    """D = A^T B with the reduction dimension as the rows of both operands (MN-major descriptors, TMA zero fill
    past the last row): the weight-gradient contraction, against torch in fp32."""
    from lossyless_b200 import _native as N

    g = torch.Generator(device=cuda_device).manual_seed(Kt + Mo)
    A = torch.randn(Kt, Mo, generator=g, device=cuda_device).to(dtype)
    B = torch.randn(Kt, No, generator=g, device=cuda_device).to(dtype)
    D = torch.full((Mo, No), float("nan"), dtype=dtype, device=cuda_device)
    N.call("ll_gemm_set_cta_pair", pair if pair else -1)
    try:
        N.call("ll_gemm_tn", N.ptr(A), N.ptr(B), N.ptr(D), Kt, Mo, No, N.dtype_code(A), N.stream())
    finally:
        N.call("ll_gemm_set_cta_pair", -1)
    ref = A.float().t() @ B.float()
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3
    assert torch.isfinite(D.float()).all()
    assert (D.float() - ref).abs().max().item() < tol * ref.abs().max().item()
    with pytest.raises(N.NativeError):
        N.call("ll_gemm_tn", N.ptr(A), N.ptr(B), N.ptr(D), Kt, Mo + 8, No, N.dtype_code(A), N.stream())


@pytest.mark.parametrize("pair", [-1, 0, 1])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("Kt,Mo,No", [(25600, 768, 3072), (6400, 3072, 768), (12800, 2304, 768), (2500, 768, 768), (137, 768, 768),
                                      (640, 128, 64)])
def test_gemm_tn_splitk(cuda_device, dtype, Kt, Mo, No, pair):
    """ll_gemm_tn_splitk: the weight-gradient contraction with the reduction split over otherwise idle SMs (float32
    partials in a caller-lent workspace, one rounding in the reduce kernel).  Against torch in fp32, against the
    single-pass ll_gemm_tn (same operands: equal up to the summation order, i.e. one 16-bit ulp), reproducible run
    to run (no atomics), and identical to the single pass when no workspace is lent."""
    from lossyless_b200 import _native as N

    if pair == 1 and (Mo % 256 or No % 128):
        pytest.skip("forced CTA pairs need 256-row / 128-column output tiles")
    g = torch.Generator(device=cuda_device).manual_seed(Kt + Mo + 1)
    A = torch.randn(Kt, Mo, generator=g, device=cuda_device).to(dtype)
    B = torch.randn(Kt, No, generator=g, device=cuda_device).to(dtype)
    ws_bytes = N.load().ll_gemm_tn_workspace_bytes(Mo, No)
    assert ws_bytes >= 4 * Mo * No
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda_device)
    D, D2, D1, D0 = (torch.full((Mo, No), float("nan"), dtype=dtype, device=cuda_device) for _ in range(4))
    N.call("ll_gemm_set_cta_pair", pair)
    try:
        N.call("ll_gemm_tn_splitk", N.ptr(A), N.ptr(B), N.ptr(D), Kt, Mo, No, N.dtype_code(A), N.ptr(ws), ws_bytes, N.stream())
        ws.fill_(0xFF)  # NaN partials: every partial the reduce reads must have been rewritten
        N.call("ll_gemm_tn_splitk", N.ptr(A), N.ptr(B), N.ptr(D2), Kt, Mo, No, N.dtype_code(A), N.ptr(ws), ws_bytes, N.stream())
        N.call("ll_gemm_tn", N.ptr(A), N.ptr(B), N.ptr(D1), Kt, Mo, No, N.dtype_code(A), N.stream())
        N.call("ll_gemm_tn_splitk", N.ptr(A), N.ptr(B), N.ptr(D0), Kt, Mo, No, N.dtype_code(A), None, 0, N.stream())
        N.call("ll_gemm_set_splitk", 0)
        D3 = torch.full_like(D, float("nan"))
        N.call("ll_gemm_tn_splitk", N.ptr(A), N.ptr(B), N.ptr(D3), Kt, Mo, No, N.dtype_code(A), N.ptr(ws), ws_bytes, N.stream())
    finally:
        N.call("ll_gemm_set_cta_pair", -1)
        N.call("ll_gemm_set_splitk", 1)
    ref = A.float().t() @ B.float()
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3
    assert torch.isfinite(D.float()).all()
    assert (D.float() - ref).abs().max().item() < tol * ref.abs().max().item()
    assert torch.equal(D, D2)
    assert torch.equal(D0, D1) and torch.equal(D3, D1)
    # float32 partials added once are at least as close to the fp32 result as the single pass
    assert (D.float() - ref).abs().mean().item() <= 1.05 * (D1.float() - ref).abs().mean().item() + 1e-6


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [1, 2, 3, 130, 301])
def test_gemm_patch_embed_reads_images_directly(cuda_device, dtype, n):
    """ll_gemm_patch_embed: CLIP's conv1 (3 -> 768, kernel = stride = 32, no bias) as a GEMM whose A operand is
    TMA-loaded from the NCHW batch through a 5-D tensor map.  Row b * 64 + gy * 8 + gx of D is patch (gy, gx) of
    image b; the padding rows (gy or gx == 7) are zero.  Reference: fp32 conv2d; and bit-equality with the im2col
    copy + plain GEMM on the same 16-bit operands.  n = 130 / 301 take the CTA-pair kernel, odd n leaves half a tile."""
    from lossyless_b200 import _native as N

    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, 3, 224, 224, generator=g).to(dtype).to(cuda_device)
    W = (torch.randn(768, 3, 32, 32, generator=g) * 0.02).to(dtype).to(cuda_device)
    D = torch.full((n * 64, 768), float("nan"), dtype=dtype, device=cuda_device)
    N.call("ll_gemm_patch_embed", N.ptr(x), N.ptr(W), N.ptr(D), n, 768, 224, 32, N.dtype_code(x), N.stream())
    D = D.reshape(n, 8, 8, 768)
    ref = torch.nn.functional.conv2d(x.float(), W.float(), stride=32).permute(0, 2, 3, 1)  # [n, 7, 7, 768]
    got = D[:, :7, :7].float()
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert (got - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
    assert float(D[:, 7].float().abs().max()) == 0.0 and float(D[:, :, 7].float().abs().max()) == 0.0
    A = x.reshape(n, 3, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5).reshape(n * 49, 3072).contiguous()
    assert torch.equal(_gemm(A, W.reshape(768, 3072)).reshape(n, 7, 7, 768), D[:, :7, :7])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(2500, 3072, 768), (25600, 3072, 768), (300, 256, 128)])
def test_gemm_quickgelu_backward_epilogue(cuda_device, dtype, M, N, K):
    """act 2: D = (A @ W^T) * QuickGELU'(u) with the pre-activation u arriving as the residual operand -- the adjoint
    of CLIP's QuickGELU fused into the data-gradient GEMM of c_proj.  Against fp32 torch autograd."""
    g = torch.Generator(device=cuda_device).manual_seed(M + N)
    A = torch.randn(M, K, generator=g, device=cuda_device).to(dtype)
    W = (torch.randn(N, K, generator=g, device=cuda_device) * K ** -0.5).to(dtype)
    u = (2.0 * torch.randn(M, N, generator=g, device=cuda_device)).to(dtype)
    D = _gemm(A, W, None, u, act=2)
    uf = u.float().requires_grad_(True)
    (uf * torch.sigmoid(1.702 * uf)).backward(A.float() @ W.float().t())
    tol = 1.5e-2 if dtype == torch.bfloat16 else 3e-3
    assert (D.float() - uf.grad).abs().max().item() < tol * uf.grad.abs().max().item()
    from lossyless_b200 import _native as Nn

    with pytest.raises(Nn.NativeError):
        _gemm(A, W, None, None, act=2)


@pytest.mark.parametrize("pair", [-1, 0, 1])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(25600, 3072, 768), (2500, 768, 3072), (6400, 768, 768), (300, 768, 2304), (7, 256, 64)])
def test_gemm_nn_reads_the_weight_as_stored(cuda_device, dtype, M, N, K, pair):
    """ll_gemm_nn: D = A @ B with B row-major [K, N] (MN-major tcgen05 operand) -- the data gradient dX = dY @ W from
    the untransposed nn.Linear weight.  Against torch in fp32, and bit-equal to the K-major GEMM over the transposed
    copy (same accumulation order); with the fused QuickGELU adjoint (act 2) as well."""
    from lossyless_b200 import _native as Nn

    if pair == 1 and (M <= 128 or N % 128):
        pytest.skip("forced CTA pairs need two row blocks and 128-column output tiles")
    g = torch.Generator(device=cuda_device).manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device=cuda_device).to(dtype)
    B = (torch.randn(K, N, generator=g, device=cuda_device) * K ** -0.5).to(dtype)
    u = (2.0 * torch.randn(M, N, generator=g, device=cuda_device)).to(dtype)
    Nn.call("ll_gemm_set_cta_pair", pair)
    try:
        D = torch.full((M, N), float("nan"), dtype=dtype, device=cuda_device)
        Nn.call("ll_gemm_nn", Nn.ptr(A), Nn.ptr(B), None, None, Nn.ptr(D), M, N, K, Nn.dtype_code(A), 0, Nn.stream())
        D2 = torch.full((M, N), float("nan"), dtype=dtype, device=cuda_device)
        Nn.call("ll_gemm_nn", Nn.ptr(A), Nn.ptr(B), None, Nn.ptr(u), Nn.ptr(D2), M, N, K, Nn.dtype_code(A), 2, Nn.stream())
    finally:
        Nn.call("ll_gemm_set_cta_pair", -1)
    ref = A.float() @ B.float()
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3
    assert torch.isfinite(D.float()).all()
    assert (D.float() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
    assert torch.equal(D, _gemm(A, B.t().contiguous()))
    assert torch.equal(D2, _gemm(A, B.t().contiguous(), None, u, act=2))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(25600, 3072, 768), (2500, 3072, 768), (300, 256, 128), (7, 64, 64)])
def test_gemm_dual_output(cuda_device, dtype, M, N, K):
    """ll_gemm_dual: U = A W^T + b and F = QuickGELU(U) from one pass.  U is bit-equal to the plain GEMM; F is
    QuickGELU of the ROUNDED U (what a separate activation pass would read), within the tanh.approx error."""
    from lossyless_b200 import _native as Nn

    g = torch.Generator(device=cuda_device).manual_seed(M + N)
    A = torch.randn(M, K, generator=g, device=cuda_device).to(dtype)
    W = (2.0 * torch.randn(N, K, generator=g, device=cuda_device) * K ** -0.5).to(dtype)
    b = torch.randn(N, generator=g, device=cuda_device)
    U = torch.full((M, N), float("nan"), dtype=dtype, device=cuda_device)
    F = torch.full((M, N), float("nan"), dtype=dtype, device=cuda_device)
    Nn.call("ll_gemm_dual", Nn.ptr(A), Nn.ptr(W), Nn.ptr(b), Nn.ptr(U), Nn.ptr(F), M, N, K, Nn.dtype_code(A), Nn.stream())
    assert torch.equal(U, _gemm(A, W, b))
    ref = U.float() * torch.sigmoid(1.702 * U.float())
    tol = 1.2e-2 if dtype == torch.bfloat16 else 2.5e-3
    assert torch.isfinite(F.float()).all()
    assert (F.float() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
