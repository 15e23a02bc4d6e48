/*
 * lossyless_b200.h -- C ABI of liblossyless_b200.so (sm_100a).
 *
 * Drop-in boundary for the encode / rate-distortion hot path of YannDubs/lossyless
 * (SURVEY.md section 8).  Plain pointers and sizes only; no torch types.  Every device
 * pointer is caller-owned, every launch is ordered on the `stream` argument
 * (a cudaStream_t passed as void*), nothing synchronises unless stated, and there is no
 * global state besides a thread-local error string.  All functions return 0 on success
 * and a negative code on failure (ll_last_error() describes it); nothing asserts.
 *
 * "replaces" cites the reference interface (file:line under /root/reference, or the
 * compressai==1.1.5 / clip==1.0 symbol the reference calls there) that each entry point
 * stands in for.  INTEGRATION.md shows the reference-side binding.
 */
#ifndef LOSSYLESS_B200_H
#define LOSSYLESS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *ll_stream_t; /* cudaStream_t */

/* element types of activations handed across the boundary */
enum { LL_F32 = 0, LL_F16 = 1, LL_BF16 = 2 };

/* error codes */
enum {
    LL_OK = 0,
    LL_ERR_INVALID = -1,  /* bad argument (null pointer, size, dtype, alignment) */
    LL_ERR_CUDA = -2,     /* CUDA runtime / driver error, see ll_last_error() */
    LL_ERR_CDF = -3,      /* pmf / CDF table does not satisfy the coder's invariants */
    LL_ERR_CORRUPT = -4,  /* container / bitstream is truncated or malformed */
    LL_ERR_ARCH = -5      /* device is not sm_100 */
};

int ll_version(void);
const char *ll_last_error(void);
/* 0 if device `dev` can run this library (compute capability 10.x), LL_ERR_ARCH otherwise */
int ll_check_device(int dev);

/* ------------------------------------------------------------------ host, load time */

/* replaces compressai._CXX.pmf_to_quantized_cdf(pmf, precision) (ops.cpp), reached from
 * EntropyBottleneck.update() at hub/compressor.py:63 and lossyless/rates.py:299,314.
 * pmf[n] -> cdf[n+1]. */
int ll_pmf_to_quantized_cdf(const float *pmf, int n, int precision, int32_t *cdf);

/* replaces compressai EntropyModel._pmf_to_cdf: per channel c,
 * cdf[c, :len+2] = pmf_to_quantized_cdf(cat(pmf[c, :pmf_length[c]], tail[c]), 16).
 * pmf [C, Lmax] row-major, tail [C], pmf_length [C]; cdf [C, Lmax+2] is zero-filled. */
int ll_build_cdf_tables(const float *pmf, const float *tail, const int32_t *pmf_length, int C,
                        int Lmax, int32_t *cdf);

/* replaces the record loop of ClipCompressor.decompress_dataset (hub/compressor.py:233-238,
 * read_uints/read_bytes :268-275): walks `n` records (">I len" + bytes) starting at
 * buf[pos] and writes offsets[n+1] = byte offset of each record's payload (offsets[n] =
 * end).  Fails with LL_ERR_CORRUPT on truncation. */
int ll_container_index(const uint8_t *buf, int64_t size, int64_t pos, int64_t n, int64_t *offsets,
                       int32_t *lens);

/* ------------------------------------------------------------------ device: rANS codec */

/* Kernel selection for the factorised-prior codec: 0 = default (encode: one stream per warp below 1024
 * rows -- latency matters -- and one stream per lane above -- throughput matters; decode: one stream per lane
 * at every row count, it wins or ties everywhere), 1 = always one stream per warp, 2 = always one stream per
 * lane.  All produce the same bytes and values (tests/test_codec_gpu.py); the knob exists for tests and
 * benchmarks.  Conditional priors (per-element table index) always take the stream-per-warp kernels.
 * Process-wide. */
int ll_rans_set_impl(int impl);

/* bytes of per-row scratch ll_eb_rans_encode needs for C symbols (worst case, escapes
 * included); always a multiple of 4. */
int ll_rans_stream_cap_bytes(int C);

/* Device-side encoder table (16 B per CDF entry: exact 64-bit reciprocal, start, freq),
 * derived from the int32 CDF table.  enc_table must hold C*cdf_stride*16 bytes.
 * replaces the per-image `self._quantized_cdf.tolist()` hand-off in compressai
 * EntropyModel.compress (hub/compressor.py:98). */
int ll_rans_build_enc_table(const int32_t *cdf, const int32_t *cdf_len, int C, int cdf_stride,
                            void *enc_table, ll_stream_t stream);

/* Fused process_z_in + quantise + rANS encode, one stream per row.
 * replaces hub/compressor.py:95,98,105-109 (process_z_in; EntropyBottleneck.compress) =
 * compressai EntropyModel.quantize("symbols", medians) + ans.RansEncoder().
 * encode_with_indexes(symbols, arange(C), cdf, cdf_len, offset), lossyless/rates.py:556-559.
 *   z        [rows, C] of z_dtype (device)       bias, exp_scale, medians  float32 [C]
 *            sym = rint((float(z) + bias) * exp_scale - medians); pass bias/exp_scale NULL
 *            to feed z_in directly (sym = rint(z - medians))
 *   scratch  [rows, cap] bytes, cap = ll_rans_stream_cap_bytes(C); row r's stream is the
 *            LAST lens[r] bytes of its slot
 *   lens     int32 [rows]        symbols_out  optional int32 [rows, C] (NULL to skip) */
int ll_eb_rans_encode(const void *z, int z_dtype, int64_t rows, int C, const float *bias,
                      const float *exp_scale, const float *medians, const void *enc_table,
                      int cdf_stride, const int32_t *cdf_len, const int32_t *offset,
                      uint8_t *scratch, int32_t *lens, int32_t *symbols_out, ll_stream_t stream);

/* Same kernel with a per-element table index and mean, for conditional priors:
 * sym = rint(z_in - means[r,c]) coded with CDF row indexes[r,c] (indexes / means NULL: channel c / medians[c]).
 * replaces compressai GaussianConditional.compress = EntropyModel.compress(inputs, indexes, means)
 * (lossyless/rates.py:705: `self.gaussian_conditional.compress(z_in, indexes, means=means_hat)`). */
int ll_rans_encode_indexed(const void *z, int z_dtype, int64_t rows, int C, const float *bias,
                           const float *exp_scale, const float *medians, const int32_t *indexes,
                           const float *means, const void *enc_table, int cdf_stride,
                           const int32_t *cdf_len, const int32_t *offset, uint8_t *scratch,
                           int32_t *lens, int32_t *symbols_out, ll_stream_t stream);

/* Compacts the per-row streams into one buffer.  mode 0: raw concatenation (what
 * ClipCompressor.compress returns, one bytes object per row, hub/compressor.py:117-119);
 * mode 1: container records ">I len" + bytes exactly as compress_dataset writes them
 * (hub/compressor.py:192-196, write_uints/write_bytes :258-265).
 *   offsets  int64 [rows+1] (device): start of each row's record in `out`
 *   out      must hold sum(lens) (+ 4*rows in mode 1) bytes; rows*(cap+4) always suffices */
int ll_rans_pack(const uint8_t *scratch, const int32_t *lens, int64_t rows, int cap, int mode,
                 uint8_t *out, int64_t *offsets, ll_stream_t stream);

/* Batched rANS decode + dequantise + process_z_out.
 * replaces hub/compressor.py:121-125 (decompress) = compressai ans.RansDecoder().
 * decode_with_indexes + EntropyModel.dequantize(medians) + process_z_out (:111-115),
 * lossyless/rates.py:561-564.
 *   data     concatenated streams (device); offsets int64 [rows+1]: stream r is
 *            data[offsets[r] : offsets[r] + lens[r]]   (lens int32 [rows])
 *   z_hat    float32 [rows, C] = (sym + medians) / exp_scale - bias  (bias/exp_scale NULL:
 *            z_hat = sym + medians)            symbols_out optional int32 [rows, C]
 *   status   int32 [rows]: 0 ok, 1 = stream shorter than the symbols require (corrupt) */
int ll_rans_decode(const uint8_t *data, const int64_t *offsets, const int32_t *lens, int64_t rows,
                   int C, const int32_t *cdf, int cdf_stride, const int32_t *cdf_len,
                   const int32_t *offset, const float *medians, const float *bias,
                   const float *exp_scale, float *z_hat, int32_t *symbols_out, int32_t *status,
                   ll_stream_t stream);

/* per-element table index / mean variant (GaussianConditional.decompress, lossyless/rates.py:716-718) */
int ll_rans_decode_indexed(const uint8_t *data, const int64_t *offsets, const int32_t *lens,
                           int64_t rows, int C, const int32_t *cdf, int cdf_stride,
                           const int32_t *cdf_len, const int32_t *offset, const float *medians,
                           const int32_t *indexes, const float *means, const float *bias,
                           const float *exp_scale, float *z_hat, int32_t *symbols_out,
                           int32_t *status, ll_stream_t stream);

/* ------------------------------------------------------- device: EntropyBottleneck */

/* Derived per-channel parameters (softplus(matrix), tanh(factor)) packed for the likelihood
 * kernels: float32 [2 * LL_EB_NPARAM, C] -- rows [0, 58) as derived (the backward kernel and
 * the gradient chain use them), rows [58, 116) the same pre-scaled by 2 log2(e) where the
 * forward kernel wants them (biases and factors times c, last-layer matrix divided by c).
 * Only filters = (3,3,3,3) (the reference's setting, hub/compressor.py:49-51,
 * config/rate/H_factorized.yaml) is supported. */
#define LL_EB_NPARAM 58
int ll_eb_pack_params(const float *const *matrices /*5*/, const float *const *biases /*5*/,
                      const float *const *factors /*4*/, int C, float *packed, ll_stream_t stream);

/* Fused rate forward: process_z_in -> (+noise | round about median) -> two
 * logits_cumulative passes -> |sigmoid diff| -> LowerBound(1e-9) -> process_z_out, plus
 * per-row rates[r] = -sum_c log(q).
 * replaces HRateFactorizedPrior.forward_help (lossyless/rates.py:534-554) = process_z_in +
 * compressai EntropyBottleneck.forward (+ _likelihood, _logits_cumulative, LowerBound) +
 * process_z_out, and hub/compressor.py:95,100-101.
 *   z [rows,C] z_dtype; noise float32 [rows,C] or NULL (eval mode: dequantise)
 *   z_hat float32 [rows,C]; q float32 [rows,C] (NULL to skip); rates float32 [rows]
 *   (NULL to skip; must be zero-filled by the caller, it is accumulated atomically) */
int ll_eb_forward(const void *z, int z_dtype, int64_t rows, int C, const float *noise,
                  const float *bias, const float *exp_scale, const float *medians,
                  const float *packed, float likelihood_bound, float *z_hat, float *q,
                  float *rates, ll_stream_t stream);

/* Backward of ll_eb_forward.  Incoming gradients (each may be NULL, not all): g_rates [rows]
 * (dL/drates), g_q [rows,C] (dL/dq, a gradient taken on the likelihood output itself, e.g.
 * -log q(s) summed outside: HRateHyperprior, lossyless/rates.py:640-650), g_zhat [rows,C]
 * (dL/dz_hat).  Writes gz float32 [rows,C] (dL/dz) and accumulates gparam float32
 * [LL_EB_NPARAM, C] (gradient w.r.t. the *derived* parameters; the softplus'/tanh' chain is
 * applied by the caller), g_bias [C], g_exp_scale [C].  noise == NULL is eval mode: v =
 * round(x - median) + median has zero derivative w.r.t. x (torch.round), so gz is zero and only
 * the parameters receive gradients; `medians` is required then.
 * replaces autograd through lossyless/rates.py:534-554 / compressai EntropyBottleneck.forward. */
int ll_eb_backward(const void *z, int z_dtype, int64_t rows, int C, const float *noise,
                   const float *bias, const float *exp_scale, const float *medians,
                   const float *packed, float likelihood_bound, const float *g_rates,
                   const float *g_q, const float *g_zhat, float *gz, float *gparam, float *g_bias,
                   float *g_exp_scale, ll_stream_t stream);

/* ------------------------------------------------------ device: Gaussian-conditional likelihood */

/* Fused compressai GaussianConditional.forward (+ per-row rate): out = x + noise (training; noise float32 [rows,C]) or
 * round(x - mean) + mean (noise == NULL); q = max(Phi((1/2 - v)/s) - Phi((-1/2 - v)/s), likelihood_bound) with
 * v = |out - mean|, s = max(scale, scale_bound); rates[r] = -sum_c log q (NULL to skip).  means may be NULL (zero
 * means), q may be NULL.  All tensors float32 [rows, C], C % 4 == 0, 16-byte aligned.
 * replaces lossyless/rates.py:638-646 (HRateHyperprior.forward_help: gaussian_conditional(...) and -log q summed). */
int ll_gc_forward(const float *x, const float *scales, const float *means, const float *noise, int64_t rows,
                  int C, float scale_bound, float likelihood_bound, float *out, float *q, float *rates,
                  ll_stream_t stream);
/* Its adjoint.  Incoming (each may be NULL, not all): g_out [rows,C], g_q [rows,C], g_rates [rows]; writes gx, gscales
 * and (when means != NULL) gmeans, all [rows,C].  Both LowerBounds pass the gradient where the input sat at or above
 * the bound or the gradient is negative (compressai ops/bound_ops.py). */
int ll_gc_backward(const float *x, const float *scales, const float *means, const float *noise, int64_t rows,
                   int C, float scale_bound, float likelihood_bound, const float *g_out, const float *g_q,
                   const float *g_rates, float *gx, float *gscales, float *gmeans, ll_stream_t stream);

/* ------------------------------------------------ device: native per-block runner (fine-tuning path) */

/* Parameters of one clip.model.ResidualAttentionBlock: 16-bit matrices as nn.Linear stores them ([out, in]), float32
 * biases and LayerNorm gains / biases. */
typedef struct {
    const void *in_proj_w, *out_proj_w, *fc_w, *cproj_w;
    const float *in_proj_b, *out_proj_b, *fc_b, *cproj_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
} ll_vit_block_params_t;

/* Where the block's parameter gradients go: the matrices' (16-bit) are overwritten, the vectors' (float32) are
 * ACCUMULATED into (zero them first). */
typedef struct {
    void *in_proj_w, *out_proj_w, *fc_w, *cproj_w;
    float *in_proj_b, *out_proj_b, *fc_b, *cproj_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
} ll_vit_block_grads_t;

/* 16-bit elements of one block's activation record [h1 | qkv | a | x1 | h2 | u | f | x2] (16 * batch * S * width),
 * the element offset of x2 (the block's output) inside it, and of the scratch one backward call needs. */
int64_t ll_vit_block_act_elems(int64_t batch, int S, int width);
int64_t ll_vit_block_x2_offset_elems(int64_t batch, int S, int width);
int64_t ll_vit_block_backward_scratch_elems(int64_t batch, int S, int width);

/* Forward of one block, x_in [batch * S, width] -> act (all activations the backward needs; x2 is the output).
 * replaces clip.model.ResidualAttentionBlock.forward under autograd (lossyless/architectures.py:351-354). */
int ll_vit_block_forward(const ll_vit_block_params_t *w, const void *x_in, void *act, int64_t batch, int S,
                         int heads, int width, int dtype, ll_stream_t stream);
/* Its adjoint: g_x holds dL/dx2 on entry and dL/dx_in on return; parameter gradients go to *g. */
int ll_vit_block_backward(const ll_vit_block_params_t *w, const void *x_in, const void *act, void *g_x,
                          void *scratch, const ll_vit_block_grads_t *g, int64_t batch, int S, int heads,
                          int width, int dtype, ll_stream_t stream);
/* The LAST block of the tower: clip.model.VisionTransformer.forward keeps x[:, 0, :] only and a block is row-wise after
 * its attention core, so the out-projection, ln_2 and the MLP -- forward, data and weight gradients -- run on the
 * `batch` class-token rows.  Same record slot and scratch size as the all-rows calls; the block's output x2_c
 * [batch, width] sits at ll_vit_block_cls_x2_offset_elems inside `act`.  Backward: g_x2_c [batch, width] is dL/dx2_c,
 * g_x [batch * S, width] receives dL/dx_in. */
int64_t ll_vit_block_cls_x2_offset_elems(int64_t batch, int S, int width);
int ll_vit_block_forward_cls(const ll_vit_block_params_t *w, const void *x_in, void *act, int64_t batch, int S,
                             int heads, int width, int dtype, ll_stream_t stream);
int ll_vit_block_backward_cls(const ll_vit_block_params_t *w, const void *x_in, const void *act, const void *g_x2_c,
                              void *g_x, void *scratch, const ll_vit_block_grads_t *g, int64_t batch, int S,
                              int heads, int width, int dtype, ll_stream_t stream);
/* Row b of a_c / x_c [batch, width] = row b * S of a / x [batch * S, width] (x[:, 0, :] of two tensors at once), and the
 * reverse into zero-filled full-size tensors.  2-byte elements, width % 8 == 0, 16-byte aligned. */
int ll_gather_cls_rows(const void *a, const void *x, void *a_c, void *x_c, int64_t batch, int S, int width, int dtype,
                       ll_stream_t stream);
int ll_scatter_cls_rows(const void *a_c, const void *x_c, void *a, void *x, int64_t batch, int S, int width, int dtype,
                        ll_stream_t stream);
/* Which fusions ll_vit_block_forward / ll_vit_block_backward use (A/B knob, tests): bit 0 = QuickGELU backward inside
 * the c_proj data-gradient GEMM's epilogue (ll_gemm act 2) instead of a separate ll_quickgelu_backward pass; bit 1 =
 * data gradients read from the untransposed weights (ll_gemm_nn) instead of four ll_transpose_pad copies per block; bit 2 =
 * QuickGELU forward as a second output of the c_fc GEMM (ll_gemm_dual) instead of a separate ll_quickgelu_forward pass.
 * Default: all on.  Returns the previous mask. */
int ll_vit_block_set_fusion(int mask);

/* ---------------------------------------------- device: flat-buffer optimiser (fine-tuning path) */

/* Everything the training step does to the unfrozen tower's parameters besides the GEMMs, over ONE flat buffer per
 * dtype.  replaces, for the encoder's parameters, torch.nn.utils.clip_grad_norm_ (config/main.yaml:113
 * gradient_clip_val), torch.optim.AdamW (config/featurizer/bottleneck_clip_lossyZ.yaml:37-47) and the autocast
 * weight casts of `precision: 16` as driven by lossyless/learnable_compressors.py:277-303.  All n are element counts,
 * multiples of 8; all buffers 16-byte aligned. */

/* floats of scratch ll_sumsq needs */
int ll_sumsq_scratch_floats(void);
/* out[0] = (accumulate ? out[0] : 0) + sum_i x[i]^2 ; x in `dtype`; fixed summation order (deterministic) */
int ll_sumsq(const void *x, int64_t n, int dtype, float *scratch, float *out, int accumulate,
             ll_stream_t stream);
/* One AdamW step (torch.optim.AdamW semantics, float32 p / m / v) with the gradient g (g_dtype) multiplied by
 * grad_scale[0] read from DEVICE memory (clip coefficient / world size; NULL = 1).  If w16 != NULL the updated
 * parameters are also written there in w16_dtype (the 16-bit compute copy: the next forward's autocast cast). */
int ll_adamw_flat(float *p, const void *g, int g_dtype, float *m, float *v, void *w16, int w16_dtype,
                  int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int64_t step, const float *grad_scale, ll_stream_t stream);
/* dst (fp16 / bf16) = src (float32) */
int ll_cast_flat(const float *src, void *dst, int dst_dtype, int64_t n, ll_stream_t stream);

/* Chain rule from ll_eb_backward's gparam [LL_EB_NPARAM, C] (gradients w.r.t. softplus(matrix), bias, tanh(factor)) to
 * the raw parameters, in their own layouts (matrices [C, f_out, f_in], biases / factors [C, f_out, 1]): one launch for
 * all 14 tensors.  m / g_m / g_b: 5 pointers, f / g_f: 4 pointers. */
int ll_eb_unpack_grads(const float *gparam, const float *const *m, const float *const *f, int C,
                       float *const *g_m, float *const *g_b, float *const *g_f, ll_stream_t stream);

/* compressai EntropyBottleneck.loss() = sum |logits_cumulative(quantiles) - target| with the parameters detached (the
 * "coder" optimiser's objective; lossyless/rates.py:426, learnable_compressors.py:293-295), forward AND the gradient
 * w.r.t. quantiles [C, 1, 3] in one launch.  `loss` (1 float) is accumulated into: zero it first.  `packed` as produced
 * by ll_eb_pack_params. */
int ll_eb_aux_loss(const float *quantiles, const float *packed, const float *target, int C, float *loss,
                   float *g_quantiles, ll_stream_t stream);

/* ------------------------------------------------------------ device: CLIP ViT-B/32 */

typedef struct ll_vit ll_vit_t;

typedef struct {
    int width, layers, heads, patch, res, out_dim; /* ViT-B/32: 768,12,12,32,224,512 */
    int dtype;                                     /* LL_F16 or LL_BF16: storage/MMA operand type */
} ll_vit_config_t;

/* Names follow clip.model.VisionTransformer's state dict.  All device pointers in
 * cfg.dtype except the LayerNorm gains/biases and linear biases (float32).  Layer arrays
 * have cfg.layers entries.  The library keeps the pointers; the caller keeps them alive. */
typedef struct {
    const void *conv1_w;   /* [width, 3*patch*patch]  (conv1.weight flattened) */
    const void *class_emb; /* [width] */
    const void *pos_emb;   /* [tokens, width] */
    const float *ln_pre_w, *ln_pre_b, *ln_post_w, *ln_post_b;
    const void *proj_t;    /* [out_dim, width] = proj.T, row-major */
    const void *const *in_proj_w;  /* [3*width, width] */
    const float *const *in_proj_b; /* [3*width] */
    const void *const *out_proj_w; /* [width, width] */
    const float *const *out_proj_b;
    const float *const *ln1_w, *const *ln1_b, *const *ln2_w, *const *ln2_b;
    const void *const *fc_w;   /* [4*width, width] */
    const float *const *fc_b;
    const void *const *cproj_w; /* [width, 4*width] */
    const float *const *cproj_b;
} ll_vit_weights_t;

/* replaces clip.load("ViT-B/32").visual (hub/compressor.py:39-40;
 * lossyless/architectures.py:367-376). */
int ll_vit_create(const ll_vit_config_t *cfg, const ll_vit_weights_t *w, ll_vit_t **out);
void ll_vit_destroy(ll_vit_t *v);
/* bytes of device workspace ll_vit_forward needs for up to `batch` images */
int64_t ll_vit_workspace_bytes(const ll_vit_t *v, int64_t batch);
/* replaces VisionTransformer.forward: `z = self.clip(X)` (hub/compressor.py:93).
 *   x [batch,3,res,res] x_dtype (LL_F16/LL_BF16/LL_F32, NCHW, device)
 *   z [batch,out_dim] in cfg.dtype (the reference's fp16 `z` before .float()) */
int ll_vit_forward(ll_vit_t *v, const void *x, int x_dtype, int64_t batch, void *z,
                   void *workspace, int64_t workspace_bytes, ll_stream_t stream);
/* number of kernel launches one ll_vit_forward of `batch` images issues */
int64_t ll_vit_launches(const ll_vit_t *v, int64_t batch);
/* max images per internal pass (default 1512 = four full rounds of 74 CTA pairs x 256 token rows);
 * a batch is split into near-equal passes */
int ll_vit_set_chunk(ll_vit_t *v, int64_t chunk);
/* ln_1 / ln_2 of every ResidualAttentionBlock (clip/model.py) are by default folded into the GEMMs either
 * side of them (ll_vit_create folds gamma/beta into library-owned copies of in_proj / c_fc; the producing
 * GEMM epilogues leave per-row statistics, see ll_gemm_ln).  on = 0 runs the separate LayerNorm kernels. */
int ll_vit_set_fuse_layernorm(ll_vit_t *v, int on);

/* 1 (default): the patch-embedding GEMM reads the image batch directly (ll_gemm_patch_embed) whenever the input has
 * the model's 16-bit dtype; 0: always go through the im2col copy (patchify kernel).  A/B and fallback knob. */
int ll_vit_set_direct_patch(ll_vit_t *v, int on);

/* im2col of CLIP's conv1 (kernel = stride = patch) with the cast to the model's 16-bit type in the same pass:
 * x [n_images, 3, res, res] (x_dtype LL_F32 / LL_F16 / LL_BF16, NCHW) -> A [n_images * (res/patch)^2, 3 * patch^2] (dtype),
 * row = patch (gy, gx) of image b, columns (channel, py, px).  The fine-tuning forward writes each view of a batch
 * straight into its slice of one A (no float32 concatenation of the views, no separate cast). */
int ll_patchify(const void *x, int x_dtype, void *A, int64_t n_images, int res, int patch, int dtype, ll_stream_t stream);

/* 1 (default): in the LAST ResidualAttentionBlock everything after the attention core runs on the class-token rows only
 * (batch rows instead of batch * tokens: the out-projection, ln_2 and the MLP are row-wise and
 * VisionTransformer.forward keeps x[:, 0, :] alone -- clip/model.py, reached from hub/compressor.py:93), which removes
 * 6 % of the encoder's FLOPs without touching z; 0: every row, as the reference computes it.  A/B knob, tests. */
int ll_vit_set_cls_last_block(ll_vit_t *v, int on);

/* Per-GEMM device timing for the roofline report: between begin and end every GEMM launch of
 * ll_vit_forward is bracketed by cudaEvents on the launch stream; end synchronises and returns
 * the summed GEMM milliseconds, algorithmic FLOPs (2*M*N*K) and launch count. */
int ll_vit_profile_begin(ll_vit_t *v);
int ll_vit_profile_end(ll_vit_t *v, double *gemm_ms, double *gemm_flops, int64_t *gemm_launches);
/* event-timed milliseconds of the last begin/end window per kernel family:
 * ms4 = {GEMM, attention, LayerNorm, patchify + token assembly} */
int ll_vit_profile_breakdown(const ll_vit_t *v, double *ms4);

/* Stand-alone entry points of the encoder's non-GEMM kernels (numerics tests).
 * ll_layernorm: h[r,:] = LN(x[r,:]) * gamma + beta, fp32 statistics, eps 1e-5 (clip LayerNorm).
 * ll_attention: qkv [batch*S, 3*H*64] (q|k|v, head-major) -> out [batch*S, H*64];
 *               softmax(q k^T / 8) v per head, S <= 64 (nn.MultiheadAttention core). */
int ll_layernorm(const void *x, int64_t rows, int width, const float *gamma, const float *beta,
                 void *h, int dtype, ll_stream_t stream);
int ll_attention(const void *qkv, void *out, int64_t batch, int S, int H, int dtype,
                 ll_stream_t stream);
/* attention implementation: 0 = mma.sync kernel, Q/K/V staged with cp.async; 1 = persistent tcgen05 kernel
 * (attention_sm100.cu); 2 = the mma.sync kernel with TMA loads / stores (fewer issued instructions; the
 * kernel is issue-bound, profiles/r01_ncu_attention_fwd_summary.txt).  Same results within rounding. */
int ll_attention_set_impl(int impl);

/* Stand-alone GEMM entry (tests, micro-benchmarks):  D[M,N] = act(A[M,K] @ W[N,K]^T + bias)
 * (+ residual), A/W/D/residual in `dtype`, bias float32 or NULL, act 0 = none,
 * 1 = QuickGELU x*sigmoid(1.702x), 2 = D = (A @ W^T + bias) * QuickGELU'(residual): the adjoint of the activation
 * fused into the data-gradient GEMM that produces its incoming gradient (residual = the pre-activation; required).
 * K % 64 == 0, N % 64 == 0. */
int ll_gemm(const void *A, const void *W, const float *bias, const void *residual, void *D,
            int64_t M, int N, int K, int dtype, int act, ll_stream_t stream);
/* GEMM with the fused-LayerNorm epilogue operands the encoder uses (nn.LayerNorm of
 * clip/model.py ResidualAttentionBlock.ln_1 / ln_2 folded into the neighbouring GEMMs):
 *   row_stats_in  float2 [stats_in_parts][M] partial (sum, sum of squares) over stat_width columns of each
 *                 A row, summed in order: every accumulator row is multiplied by rsqrt(var + 1e-5) before
 *                 the bias.  With W row-centred and pre-scaled by gamma and bias' = bias + W beta (host,
 *                 load time) this is LN(A) W^T + bias.
 *   row_stats_out float2 [N / 64][M]: the same partials of the D rows as stored, one part per 64 columns;
 *                 one writer per slot and a tile-independent order, so results are reproducible.
 * Either may be NULL (both NULL == ll_gemm). */
int ll_gemm_ln(const void *A, const void *W, const float *bias, const void *residual, void *D,
               int64_t M, int N, int K, int dtype, int act, const void *row_stats_in,
               int stats_in_parts, void *row_stats_out, int stat_width, ll_stream_t stream);
/* Patch embedding without im2col (CLIP conv1: 3 -> width, kernel = stride = patch, no bias; clip/model.py
 * VisionTransformer.conv1, reached from hub/compressor.py:93): D[n * 64, N] = patches(x) @ W[N, 3 patch^2]^T with the A
 * operand TMA-loaded straight from the image batch x [n, 3, res, res] (16-bit `dtype`) through a 5-D tensor map.
 * Needs patch == 32 and res / patch <= 8.  Every image owns 64 rows of D: row gy * 8 + gx holds patch (gy, gx); the
 * padding rows (gy or gx >= res / patch) are zero. */
int ll_gemm_patch_embed(const void *x, const void *W, void *D, int64_t n_images, int N, int res,
                        int patch, int dtype, ll_stream_t stream);

/* D[Mo, No] = A[Kt, Mo]^T @ B[Kt, No] on the same tcgen05 kernel with MN-major operand descriptors: the
 * weight-gradient contraction dW = dY^T X of a linear layer (torch autograd through nn.Linear in the reference's
 * unfrozen CLIP tower, lossyless/architectures.py:351-376) taken straight from the row-major activations and
 * gradients -- no transposed copies.  A, B, D in `dtype` (LL_F16 / LL_BF16), fp32 accumulation.
 * Mo % 128 == 0, No % 64 == 0, Kt arbitrary (rows past Kt are zero-filled by TMA). */
int ll_gemm_tn(const void *A, const void *B, void *D, int64_t Kt, int Mo, int No, int dtype, ll_stream_t stream);

/* U[M,N] = A[M,K] @ W[N,K]^T + bias and F = QuickGELU(U) (both in `dtype`; F computed from the rounded U) from ONE GEMM
 * pass: c_fc of the fine-tuning forward keeps the pre-activation for the backward pass and feeds the activation to c_proj
 * -- replaces `u = c_fc(h); f = QuickGELU(u)` of clip.model.ResidualAttentionBlock.mlp under autograd. */
int ll_gemm_dual(const void *A, const void *W, const float *bias, void *U, void *F, int64_t M, int N, int K, int dtype,
                 ll_stream_t stream);

/* D[M,N] = act(A[M,K] @ B[K,N] + bias) (+ residual) with B ROW-MAJOR [K, N] (an MN-major tcgen05 operand: 64 x 64 TMA
 * boxes, transpose bit of the instruction descriptor): the data gradient dX = dY @ W of a linear layer straight from the
 * weight as nn.Linear stores it ([out, in] = [K, N]) -- replaces autograd's `grad_output.matmul(weight)` without a
 * transposed weight copy.  Same epilogue options as ll_gemm (act 0 / 1 / 2).  K % 64 == 0, N % 64 == 0. */
int ll_gemm_nn(const void *A, const void *B, const float *bias, const void *residual, void *D, int64_t M, int N, int K,
               int dtype, int act, ll_stream_t stream);

/* ll_gemm_tn with split-K: when the output is only a fraction of a round of tiles (a [768..3072]^2 weight gradient is
 * 36-144 tiles for a 25 600-deep reduction) the reduction is split 2-4 ways over otherwise idle SMs, every split leaves a
 * float32 partial in `workspace` (ll_gemm_tn_workspace_bytes(Mo, No) bytes, 16-byte aligned) and a second kernel adds
 * them with one rounding to the 16-bit result.  Falls back to ll_gemm_tn's single pass when splitting does not help.
 * ll_gemm_set_splitk(0) switches the splitting off (A/B knob). */
int64_t ll_gemm_tn_workspace_bytes(int Mo, int No);
int ll_gemm_tn_splitk(const void *A, const void *B, void *D, int64_t Kt, int Mo, int No, int dtype,
                      void *workspace, int64_t workspace_bytes, ll_stream_t stream);
int ll_gemm_set_splitk(int on);
/* Force the GEMM's tile width (64/128/192/256; 0 = automatic wave-aware choice). Tests, tuning. */
int ll_gemm_set_tile_n(int bn);
/* CTA-pair (tcgen05 cta_group::2, 256-row tiles) policy: -1 automatic, 0 never, 1 always. */
int ll_gemm_set_cta_pair(int mode);

/* ----------------------------------------- device: ViT encoder backward (fine-tuning) */

/* The reference fine-tunes through `clip.load(...).visual.float()` with torch autograd under AMP
 * (lossyless/architectures.py:351-376, lossyless/learnable_compressors.py:277-303).  Here every dense
 * contraction of the backward pass (dX = dY W, dW = dY^T X) is an ll_gemm call on transposed 16-bit
 * copies; the entry points below are the rest of that backward pass (lossyless_b200/clip_vit.py drives
 * them).  All tensors device, row-major, `dtype` LL_F16 or LL_BF16 unless stated. */

/* dst [C, Rpad] = src [R, C]^T, columns R..Rpad-1 zero (2-byte elements; Rpad >= R).  The token count
 * becomes the K of a weight-gradient GEMM and must be a multiple of 64. */
int ll_transpose_pad(const void *src, void *dst, int64_t R, int C, int64_t Rpad, ll_stream_t stream);
/* the same plus colsum[c] += sum_r src[r, c] (float32; the bias gradient of the tensor being transposed for
 * its weight-gradient GEMM, so it is read once).  C and Rpad even, `dtype` LL_F16 / LL_BF16. */
int ll_transpose_pad_colsum(const void *src, void *dst, int64_t R, int C, int64_t Rpad, int dtype, float *colsum,
                            ll_stream_t stream);
/* out[c] += sum_r src[r, c] (float32 accumulate: bias gradients).  out must be initialised. */
int ll_colsum(const void *src, int64_t R, int C, int dtype, float *out, ll_stream_t stream);
/* VisionTransformer.forward's token assembly for the training path (the inference path fuses it with ln_pre inside
 * ll_vit_forward): tok [B, S, W] = cat([class_embedding, patch [B, S - 1, W]], dim = 1) + positional_embedding [S, W]
 * (the reference's clip VisionTransformer.forward; one 16-bit rounding per element, as torch's 16-bit add).  W % 8 == 0. */
int ll_embed_tokens(const void *patch, const void *cls, const void *pos, void *tok, int64_t B, int S, int W, int dtype,
                    ll_stream_t stream);
/* its adjoint, g_tok read once: g_patch [B, S - 1, W] = g_tok[:, 1:, :] (contiguous: conv1's weight-gradient operand),
 * g_pos [S, W] += sum_b g_tok[b], g_cls [W] += sum_b g_tok[b, 0] (float32; both must be initialised). */
int ll_embed_backward(const void *g_tok, void *g_patch, float *g_pos, float *g_cls, int64_t B, int S, int W, int dtype,
                      ll_stream_t stream);
/* clip/model.py QuickGELU: f = u * sigmoid(1.702 u);  g_u = g_f * d f / d u.  n elements. */
int ll_quickgelu_forward(const void *u, void *f, int64_t n, int dtype, ll_stream_t stream);
int ll_quickgelu_backward(const void *u, const void *g_f, void *g_u, int64_t n, int dtype, ll_stream_t stream);
/* Backward of ll_layernorm (clip/model.py LayerNorm, eps 1e-5), statistics recomputed from x:
 *   g_x = (g_res or 0) + dLN(x)^T g_h;   g_gamma, g_beta float32 [width], ACCUMULATED (atomics).
 * width in {512, 768, 1024}; g_res may be NULL; g_x must not alias the inputs. */
int ll_layernorm_backward(const void *x, const void *g_h, const float *gamma, const void *g_res, void *g_x,
                          float *g_gamma, float *g_beta, int64_t rows, int width, int dtype,
                          ll_stream_t stream);
/* Backward of ll_attention: qkv [batch*S, 3*H*64], g_out [batch*S, H*64] -> g_qkv [batch*S, 3*H*64].
 * The probabilities are recomputed from qkv.  S <= 64. */
int ll_attention_backward(const void *qkv, const void *g_out, void *g_qkv, int64_t batch, int S, int H,
                          int dtype, ll_stream_t stream);
/* 0 (default) = mma.sync tensor-core kernel fed by ldmatrix / ldmatrix.trans from one orientation of each
 * operand (P and dS rounded to `dtype` between the products, as flash-attention backward does), 1 = fp32
 * CUDA-core kernel, 2 = the first tensor-core version (operands copied in both orientations). */
int ll_attention_backward_set_impl(int impl);

/* ------------------------------------------------------ device: CLIP preprocessing */

/* uint8 images -> bicubic resize (Pillow's 8-bit ImagingResample, bit-exact) -> centre crop n_px ->
 * /255 -> (x - mean) / std -> [B, 3, n_px, n_px] in out_dtype.
 * replaces the CPU transform the reference's DataLoader applies before compress_dataset
 * (clip `_transform` returned by clip.load, hub/compressor.py:39; utils/data/images.py:383-389).
 *   img      uint8, layout 0: [B, H, W, 3] (PIL / numpy), layout 1: [B, 3, H, W]          (device)
 *   coef_h   int32 [n_px, ksize_h], bounds_h int32 [n_px, 2] = (first input column, count) per output
 *            column of the crop window; coef_v / bounds_v likewise for rows: Pillow's 22-bit
 *            fixed-point coefficients, computed by the caller (lossyless_b200/preprocess.py)   (device)
 *   mean3, std3  float[3] on the HOST;   tmp uint8 [B, H, n_px, 3] scratch                  (device) */
int ll_clip_preprocess(const uint8_t *img, int64_t B, int H, int W, int layout, const int32_t *coef_h,
                       const int32_t *bounds_h, int ksize_h, const int32_t *coef_v,
                       const int32_t *bounds_v, int ksize_v, const float *mean3, const float *std3,
                       int n_px, uint8_t *tmp, void *out, int out_dtype, ll_stream_t stream);
/* 0 (default): images whose pixels + intermediate fit shared memory take a fused one-CTA-per-image kernel;
 * 1: always the general two-pass kernels.  Same bits either way (tests/test_preprocess_gpu.py). */
int ll_clip_preprocess_set_impl(int impl);

/* ----------------------------------------------------------- device: distortions */

/* Row-wise L2 normalisation (F.normalize, eps 1e-12) + zs_norm @ zs_norm^T tile, fp32.
 * replaces lossyless/distortions.py:319-324.
 *   zs float32 [n, d] (n = 2B) -> zs_norm [n,d], inv_norm [n], logits [n, ld] (tile written
 *   at column 0; ld >= n lets the caller gather remote tiles to its right) */
int ll_contrastive_logits(const float *zs, int n, int d, float *zs_norm, float *inv_norm,
                          float *logits, int ld, ll_stream_t stream);
/* backward of the above: g_logits [n, ld_g] (local tile) -> g_zs [n, d]; d <= 256. */
int ll_contrastive_logits_backward(const float *g_logits, int ld_g, const float *zs_norm,
                                   const float *inv_norm, int n, int d, float *g_zs,
                                   ll_stream_t stream);

/* InfoNCE over the gathered logits without materialising the masked copy.
 * replaces ContrastiveDistortion.compute_loss (lossyless/distortions.py:340-405).
 *   logits [n, n_cols] (own tile first), pos_add = -log(to_mult) or 0, inv_temp = 1/temperature
 *   loss [n] = cross_entropy(masked(logits)/T, pos_idx);  lse [n] saved for backward */
int ll_contrastive_ce_forward(const float *logits, int n, int n_cols, int ld, float pos_add,
                              float inv_temp, float *loss, float *lse, ll_stream_t stream);
/* grad wrt the LOCAL tile only (lossyless/helpers.py:885-890) and wrt inv_temp. */
int ll_contrastive_ce_backward(const float *logits, int n, int n_cols, int ld, float pos_add,
                               float inv_temp, const float *lse, const float *g_loss,
                               float *g_logits_local, int ld_g, float *g_inv_temp,
                               ll_stream_t stream);

/* The same two entry points with 1/temperature read from DEVICE memory (one float): the temperature is a trained
 * parameter (lossyless/distortions.py:228-236), and passing it by value forces the host to wait for the GPU in
 * the middle of every training step. */
int ll_contrastive_ce_forward_dev(const float *logits, int n, int n_cols, int ld, float pos_add,
                                  const float *inv_temp_dev, float *loss, float *lse, ll_stream_t stream);
int ll_contrastive_ce_backward_dev(const float *logits, int n, int n_cols, int ld, float pos_add,
                                   const float *inv_temp_dev, const float *lse, const float *g_loss,
                                   float *g_logits_local, int ld_g, float *g_inv_temp, ll_stream_t stream);

/* replaces LossyZDistortion.forward (lossyless/distortions.py:445):
 * dist[r] = sum_c |z_hat - z + 1e-6|;  optional backward sign tensor. */
int ll_lossy_z_forward(const float *z_hat, const float *z, int64_t rows, int C, float eps,
                       float *dist, ll_stream_t stream);
int ll_lossy_z_backward(const float *z_hat, const float *z, int64_t rows, int C, float eps,
                        const float *g_dist, float *g_zhat, ll_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LOSSYLESS_B200_H */
