// Native per-block runner of the fine-tuning path (sm_100a): ONE call sequences the forward pass of a CLIP
// ResidualAttentionBlock (keeping the activations its backward needs), and ONE call its hand-written adjoint,
// instead of ~13 / ~26 Python-issued launches with their tensor allocations in between.  With the flat training state
// the training step was bound by the HOST (746 launches + torch bookkeeping per 27 ms step, GPU kernels summing to 24 ms):
// a step is now ~100 host calls.
//
// Replaces torch autograd through clip.model.ResidualAttentionBlock.forward (x += attn(ln_1(x)); x += mlp(ln_2(x))) as
// reached from lossyless/architectures.py:351-354 under lossyless/learnable_compressors.py:277-303.  Every dense
// contraction (forward, data gradient, weight gradient) is the tcgen05 GEMM of gemm_sm100.cu; the rest are the kernels
// of vit.cu / vit_train.cu.  The caller (clip_vit._VitTrainFn) owns all memory and launches the data-parallel
// all-reduce of a block's gradients between two backward calls.
//
// Activation record of one block, 16-bit, M = batch * tokens rows each, in this order:
//   h1 [M,W] ln_1(x) | qkv [M,3W] | a [M,W] attention output | x1 [M,W] | h2 [M,W] ln_2(x1) | u [M,4W] | f [M,4W] QuickGELU(u)
//   | x2 [M,W] block output
#include "gemm_sm100.cuh"

using namespace ll;

namespace {

struct ActView {
    uint8_t *h1, *qkv, *a, *x1, *h2, *u, *f, *x2;
};

inline ActView act_view(void *act, int64_t M, int W)
{
    uint8_t *p = static_cast<uint8_t *>(act);
    ActView v;
    const int64_t row = 2;  // bytes per element
    v.h1 = p, p += M * W * row;
    v.qkv = p, p += M * 3 * W * row;
    v.a = p, p += M * W * row;
    v.x1 = p, p += M * W * row;
    v.h2 = p, p += M * W * row;
    v.u = p, p += M * 4 * W * row;
    v.f = p, p += M * 4 * W * row;
    v.x2 = p;
    return v;
}

int g_fusion = 0x7fffffff;  // ll_vit_block_set_fusion

}  // namespace

extern "C" int ll_vit_block_set_fusion(int mask)
{
    const int old = g_fusion;
    g_fusion = mask;
    return old;
}

extern "C" int64_t ll_vit_block_act_elems(int64_t batch, int S, int width) { return batch * S * (int64_t)width * 16; }

extern "C" int64_t ll_vit_block_x2_offset_elems(int64_t batch, int S, int width) { return batch * S * (int64_t)width * 15; }

// scratch of one backward call: g_f / g_u [M,4W], g_h2 [M,W], g_x1 [M,W], g_a [M,W], g_qkv [M,3W], g_h1 [M,W], then the four
// transposed weight copies (12 W^2 elements), then the split-K workspace of the weight-gradient GEMMs (4 float32
// partials of the largest gradient, 4 W^2 elements: 32 W^2 16-bit elements)
extern "C" int64_t ll_vit_block_backward_scratch_elems(int64_t batch, int S, int width)
{
    return batch * S * (int64_t)width * 11 + 12LL * width * width + 32LL * width * width;
}

extern "C" int ll_vit_block_forward(const ll_vit_block_params_t *w, const void *x_in, void *act, int64_t batch, int S, int heads,
                                    int width, int dtype, ll_stream_t stream)
{
    LL_REQUIRE(w && x_in && act && batch > 0 && S > 0 && heads > 0 && width > 0, "vit_block_forward: bad arguments");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "vit_block_forward: dtype must be LL_F16 or LL_BF16");
    const int64_t M = batch * S;
    const int W = width;
    const ActView v = act_view(act, M, W);
    int rc;
    if ((rc = ll_layernorm(x_in, M, W, w->ln1_w, w->ln1_b, v.h1, dtype, stream))) return rc;
    if ((rc = ll_gemm(v.h1, w->in_proj_w, w->in_proj_b, nullptr, v.qkv, M, 3 * W, W, dtype, 0, stream))) return rc;
    if ((rc = ll_attention(v.qkv, v.a, batch, S, heads, dtype, stream))) return rc;
    if ((rc = ll_gemm(v.a, w->out_proj_w, w->out_proj_b, x_in, v.x1, M, W, W, dtype, 0, stream))) return rc;
    if ((rc = ll_layernorm(v.x1, M, W, w->ln2_w, w->ln2_b, v.h2, dtype, stream))) return rc;
    if (g_fusion & 4) {  // u and f = QuickGELU(u) from one GEMM pass
        if ((rc = ll_gemm_dual(v.h2, w->fc_w, w->fc_b, v.u, v.f, M, 4 * W, W, dtype, stream))) return rc;
    } else {
        if ((rc = ll_gemm(v.h2, w->fc_w, w->fc_b, nullptr, v.u, M, 4 * W, W, dtype, 0, stream))) return rc;
        if ((rc = ll_quickgelu_forward(v.u, v.f, M * 4 * (int64_t)W, dtype, stream))) return rc;
    }
    if ((rc = ll_gemm(v.f, w->cproj_w, w->cproj_b, v.x1, v.x2, M, W, 4 * W, dtype, 0, stream))) return rc;
    return LL_OK;
}

extern "C" int ll_vit_block_backward(const ll_vit_block_params_t *w, const void *x_in, const void *act, void *g_x, void *scratch,
                                     const ll_vit_block_grads_t *g, int64_t batch, int S, int heads, int width, int dtype,
                                     ll_stream_t stream)
{
    LL_REQUIRE(w && x_in && act && g_x && scratch && g && batch > 0 && S > 0, "vit_block_backward: bad arguments");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "vit_block_backward: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(width % 64 == 0, "vit_block_backward: width must be a multiple of 64");
    const int64_t M = batch * S;
    const int W = width;
    const ActView v = act_view(const_cast<void *>(act), M, W);
    uint8_t *p = static_cast<uint8_t *>(scratch);
    uint8_t *g_u = p;     p += M * 4 * W * 2;  // holds g_f first, g_u in place after the QuickGELU backward
    uint8_t *g_h2 = p;    p += M * W * 2;
    uint8_t *g_x1 = p;    p += M * W * 2;
    uint8_t *g_a = p;     p += M * W * 2;
    uint8_t *g_qkv = p;   p += M * 3 * W * 2;
    uint8_t *g_h1 = p;    p += M * W * 2;
    // transposed 16-bit weight copies: dX = dY @ W is the K-major GEMM D = A Wt^T with Wt = W^T [in, out]
    uint8_t *cproj_t = p; p += (int64_t)4 * W * W * 2;
    uint8_t *fc_t = p;    p += (int64_t)4 * W * W * 2;
    uint8_t *out_t = p;   p += (int64_t)W * W * 2;
    uint8_t *in_t = p;    p += (int64_t)3 * W * W * 2;
    void *ws = p;  // split-K workspace (16-byte aligned: every piece above is a multiple of 128 bytes)
    const int64_t ws_bytes = 64LL * W * W;
    int rc;
    // data gradients dX = dY @ W: from the weight as stored ([out, in] = [K, N] row-major, ll_gemm_nn), or -- fusion bit 1
    // off -- as the K-major GEMM D = A Wt^T over a transposed copy Wt = W^T [in, out] made here
    const bool nn = (g_fusion & 2) != 0;
    auto dgrad = [&](const void *dY, const void *Wsrc, const void *Wt, const void *res, void *dX, int64_t m, int n_in, int k_out, int act) {
        return nn ? ll_gemm_nn(dY, Wsrc, nullptr, res, dX, m, n_in, k_out, dtype, act, stream)
                  : ll_gemm(dY, Wt, nullptr, res, dX, m, n_in, k_out, dtype, act, stream);
    };
    if (!nn) {
        // W [out, in] -> Wt [in, out]  (transpose_pad(src [R, C]) -> [C, Rpad]; R = out is a multiple of 64 here)
        if ((rc = ll_transpose_pad(w->cproj_w, cproj_t, W, 4 * W, W, stream))) return rc;
        if ((rc = ll_transpose_pad(w->fc_w, fc_t, 4 * W, W, 4 * W, stream))) return rc;
        if ((rc = ll_transpose_pad(w->out_proj_w, out_t, W, W, W, stream))) return rc;
        if ((rc = ll_transpose_pad(w->in_proj_w, in_t, 3 * W, W, 3 * W, stream))) return rc;
    }
    // x2 = x1 + f @ c_proj^T + b
    // g_u = (g_x @ c_proj) * QuickGELU'(u): the activation's adjoint rides in the GEMM's epilogue (u arrives as the
    // "residual" tile); unfused: g_f first, then a separate elementwise pass in place
    const bool fuse_gelu_bwd = (g_fusion & 1) != 0;
    if ((rc = dgrad(g_x, w->cproj_w, cproj_t, fuse_gelu_bwd ? v.u : nullptr, g_u, M, 4 * W, W, fuse_gelu_bwd ? 2 : 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_x, v.f, g->cproj_w, M, W, 4 * W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_x, M, W, dtype, g->cproj_b, stream))) return rc;
    if (!fuse_gelu_bwd && (rc = ll_quickgelu_backward(v.u, g_u, g_u, M * 4 * (int64_t)W, dtype, stream))) return rc;  // in place
    if ((rc = dgrad(g_u, w->fc_w, fc_t, nullptr, g_h2, M, W, 4 * W, 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_u, v.h2, g->fc_w, M, 4 * W, W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_u, M, 4 * W, dtype, g->fc_b, stream))) return rc;
    if ((rc = ll_layernorm_backward(v.x1, g_h2, w->ln2_w, g_x, g_x1, g->ln2_w, g->ln2_b, M, W, dtype, stream))) return rc;
    // x1 = x + a @ out_proj^T + b
    if ((rc = dgrad(g_x1, w->out_proj_w, out_t, nullptr, g_a, M, W, W, 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_x1, v.a, g->out_proj_w, M, W, W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_x1, M, W, dtype, g->out_proj_b, stream))) return rc;
    if ((rc = ll_attention_backward(v.qkv, g_a, g_qkv, batch, S, heads, dtype, stream))) return rc;
    if ((rc = dgrad(g_qkv, w->in_proj_w, in_t, nullptr, g_h1, M, W, 3 * W, 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_qkv, v.h1, g->in_proj_w, M, 3 * W, W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_qkv, M, 3 * W, dtype, g->in_proj_b, stream))) return rc;
    // dL/dx_in = LN1 backward + the residual branch (g_x1); overwrites g_x, whose old value is dead by now
    if ((rc = ll_layernorm_backward(x_in, g_h1, w->ln1_w, g_x1, g_x, g->ln1_w, g->ln1_b, M, W, dtype, stream))) return rc;
    return LL_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// The LAST block of the tower.  VisionTransformer.forward keeps x[:, 0, :] only, and after the attention core a
// ResidualAttentionBlock is row-wise: the out-projection, ln_2 and the MLP of the last block -- forward, data gradients
// and weight gradients -- run on the `batch` class-token rows instead of batch * S (the other rows' outputs are never
// read and their gradients are exactly zero).  Same function, 6 % fewer FLOPs per training step.
//
// Activation record (inside the same ll_vit_block_act_elems slot): h1 [M,W] | qkv [M,3W] | a [M,W] as before, then the
// class-token tensors a_c [B,W] | x_c [B,W] | x1_c [B,W] | h2_c [B,W] | u_c [B,4W] | f_c [B,4W] | x2_c [B,W].
namespace {

struct ClsView {
    uint8_t *h1, *qkv, *a, *a_c, *x_c, *x1_c, *h2_c, *u_c, *f_c, *x2_c;
};

inline ClsView cls_view(void *act, int64_t M, int64_t B, int W)
{
    uint8_t *p = static_cast<uint8_t *>(act);
    ClsView v;
    v.h1 = p, p += M * W * 2;
    v.qkv = p, p += M * 3 * W * 2;
    v.a = p, p += M * W * 2;
    v.a_c = p, p += B * W * 2;
    v.x_c = p, p += B * W * 2;
    v.x1_c = p, p += B * W * 2;
    v.h2_c = p, p += B * W * 2;
    v.u_c = p, p += B * 4 * W * 2;
    v.f_c = p, p += B * 4 * W * 2;
    v.x2_c = p;
    return v;
}

}  // namespace

extern "C" int64_t ll_vit_block_cls_x2_offset_elems(int64_t batch, int S, int width)
{
    return batch * S * (int64_t)width * 5 + batch * (int64_t)width * 12;
}

extern "C" int ll_vit_block_forward_cls(const ll_vit_block_params_t *w, const void *x_in, void *act, int64_t batch, int S, int heads,
                                        int width, int dtype, ll_stream_t stream)
{
    LL_REQUIRE(w && x_in && act && batch > 0 && S > 0 && heads > 0 && width > 0, "vit_block_forward_cls: bad arguments");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "vit_block_forward_cls: dtype must be LL_F16 or LL_BF16");
    const int64_t M = batch * S, B = batch;
    const int W = width;
    const ClsView v = cls_view(act, M, B, W);
    int rc;
    if ((rc = ll_layernorm(x_in, M, W, w->ln1_w, w->ln1_b, v.h1, dtype, stream))) return rc;
    if ((rc = ll_gemm(v.h1, w->in_proj_w, w->in_proj_b, nullptr, v.qkv, M, 3 * W, W, dtype, 0, stream))) return rc;
    if ((rc = ll_attention(v.qkv, v.a, batch, S, heads, dtype, stream))) return rc;
    if ((rc = ll_gather_cls_rows(v.a, x_in, v.a_c, v.x_c, batch, S, W, dtype, stream))) return rc;
    if ((rc = ll_gemm(v.a_c, w->out_proj_w, w->out_proj_b, v.x_c, v.x1_c, B, W, W, dtype, 0, stream))) return rc;
    if ((rc = ll_layernorm(v.x1_c, B, W, w->ln2_w, w->ln2_b, v.h2_c, dtype, stream))) return rc;
    if (g_fusion & 4) {
        if ((rc = ll_gemm_dual(v.h2_c, w->fc_w, w->fc_b, v.u_c, v.f_c, B, 4 * W, W, dtype, stream))) return rc;
    } else {
        if ((rc = ll_gemm(v.h2_c, w->fc_w, w->fc_b, nullptr, v.u_c, B, 4 * W, W, dtype, 0, stream))) return rc;
        if ((rc = ll_quickgelu_forward(v.u_c, v.f_c, B * 4 * (int64_t)W, dtype, stream))) return rc;
    }
    if ((rc = ll_gemm(v.f_c, w->cproj_w, w->cproj_b, v.x1_c, v.x2_c, B, W, 4 * W, dtype, 0, stream))) return rc;
    return LL_OK;
}

extern "C" int ll_vit_block_backward_cls(const ll_vit_block_params_t *w, const void *x_in, const void *act, const void *g_x2_c, void *g_x,
                                         void *scratch, const ll_vit_block_grads_t *g, int64_t batch, int S, int heads, int width,
                                         int dtype, ll_stream_t stream)
{
    LL_REQUIRE(w && x_in && act && g_x2_c && g_x && scratch && g && batch > 0 && S > 0, "vit_block_backward_cls: bad arguments");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "vit_block_backward_cls: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(width % 64 == 0, "vit_block_backward_cls: width must be a multiple of 64");
    const int64_t M = batch * S, B = batch;
    const int W = width;
    const ClsView v = cls_view(const_cast<void *>(act), M, B, W);
    // scratch (the size ll_vit_block_backward_scratch_elems reports covers it): class-token tensors, the two scattered
    // full-size gradients, g_qkv, g_h1, the transposed weights, the split-K workspace
    uint8_t *p = static_cast<uint8_t *>(scratch);
    uint8_t *g_u_c = p;   p += B * 4 * W * 2;
    uint8_t *g_h2_c = p;  p += B * W * 2;
    uint8_t *g_x1_c = p;  p += B * W * 2;
    uint8_t *g_a_c = p;   p += B * W * 2;
    p = static_cast<uint8_t *>(scratch) + ((B * 7 * W * 2 + 127) / 128) * 128;
    uint8_t *g_a = p;     p += M * W * 2;
    uint8_t *g_x1 = p;    p += M * W * 2;
    uint8_t *g_qkv = p;   p += M * 3 * W * 2;
    uint8_t *g_h1 = p;    p += M * W * 2;
    uint8_t *cproj_t = p; p += (int64_t)4 * W * W * 2;
    uint8_t *fc_t = p;    p += (int64_t)4 * W * W * 2;
    uint8_t *out_t = p;   p += (int64_t)W * W * 2;
    uint8_t *in_t = p;    p += (int64_t)3 * W * W * 2;
    void *ws = p;
    const int64_t ws_bytes = 64LL * W * W;
    int rc;
    const bool nn = (g_fusion & 2) != 0;
    auto dgrad = [&](const void *dY, const void *Wsrc, const void *Wt, const void *res, void *dX, int64_t m, int n_in, int k_out, int act) {
        return nn ? ll_gemm_nn(dY, Wsrc, nullptr, res, dX, m, n_in, k_out, dtype, act, stream)
                  : ll_gemm(dY, Wt, nullptr, res, dX, m, n_in, k_out, dtype, act, stream);
    };
    if (!nn) {
        if ((rc = ll_transpose_pad(w->cproj_w, cproj_t, W, 4 * W, W, stream))) return rc;
        if ((rc = ll_transpose_pad(w->fc_w, fc_t, 4 * W, W, 4 * W, stream))) return rc;
        if ((rc = ll_transpose_pad(w->out_proj_w, out_t, W, W, W, stream))) return rc;
        if ((rc = ll_transpose_pad(w->in_proj_w, in_t, 3 * W, W, 3 * W, stream))) return rc;
    }
    // x2_c = x1_c + f_c @ c_proj^T + b
    const bool fuse_gelu_bwd = (g_fusion & 1) != 0;
    if ((rc = dgrad(g_x2_c, w->cproj_w, cproj_t, fuse_gelu_bwd ? v.u_c : nullptr, g_u_c, B, 4 * W, W, fuse_gelu_bwd ? 2 : 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_x2_c, v.f_c, g->cproj_w, B, W, 4 * W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_x2_c, B, W, dtype, g->cproj_b, stream))) return rc;
    if (!fuse_gelu_bwd && (rc = ll_quickgelu_backward(v.u_c, g_u_c, g_u_c, B * 4 * (int64_t)W, dtype, stream))) return rc;
    if ((rc = dgrad(g_u_c, w->fc_w, fc_t, nullptr, g_h2_c, B, W, 4 * W, 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_u_c, v.h2_c, g->fc_w, B, 4 * W, W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_u_c, B, 4 * W, dtype, g->fc_b, stream))) return rc;
    if ((rc = ll_layernorm_backward(v.x1_c, g_h2_c, w->ln2_w, g_x2_c, g_x1_c, g->ln2_w, g->ln2_b, B, W, dtype, stream))) return rc;
    // x1_c = x_c + a_c @ out_proj^T + b
    if ((rc = dgrad(g_x1_c, w->out_proj_w, out_t, nullptr, g_a_c, B, W, W, 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_x1_c, v.a_c, g->out_proj_w, B, W, W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_x1_c, B, W, dtype, g->out_proj_b, stream))) return rc;
    // back to all rows: dL/da and the residual branch are zero outside the class-token rows
    if ((rc = ll_scatter_cls_rows(g_a_c, g_x1_c, g_a, g_x1, batch, S, W, dtype, stream))) return rc;
    if ((rc = ll_attention_backward(v.qkv, g_a, g_qkv, batch, S, heads, dtype, stream))) return rc;
    if ((rc = dgrad(g_qkv, w->in_proj_w, in_t, nullptr, g_h1, M, W, 3 * W, 0))) return rc;
    if ((rc = ll_gemm_tn_splitk(g_qkv, v.h1, g->in_proj_w, M, 3 * W, W, dtype, ws, ws_bytes, stream))) return rc;
    if ((rc = ll_colsum(g_qkv, M, 3 * W, dtype, g->in_proj_b, stream))) return rc;
    if ((rc = ll_layernorm_backward(x_in, g_h1, w->ln1_w, g_x1, g_x, g->ln1_w, g->ln1_b, M, W, dtype, stream))) return rc;
    return LL_OK;
}
