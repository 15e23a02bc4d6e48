// Backward-pass kernels of the CLIP ViT-B/32 encoder for rate-distortion fine-tuning with the encoder
// unfrozen (the reference trains through `clip.load(...).visual.float()` with torch autograd under AMP:
// lossyless/architectures.py:351-376, lossyless/learnable_compressors.py:277-303).  Every dense
// contraction of the backward pass -- data gradients dY W and weight gradients dY^T X -- runs on the same
// tcgen05 GEMM as the forward pass (csrc/gemm_sm100.cu computes D = A W^T with both operands K-major), fed
// with transposed 16-bit copies made by transpose_pad_kernel.  This file holds what is not a GEMM:
//   transpose_pad     [R, C] -> [C, Rpad] (zero padded: the token count becomes the GEMM's K, K % 64 == 0)
//   colsum            bias gradients (fp32 column sums)
//   quickgelu fwd/bwd x * sigmoid(1.702 x) on a saved pre-activation (clip/model.py QuickGELU)
//   layernorm_backward  fp32 statistics recomputed from the saved input; gamma/beta gradients by atomics
//   attention_backward  softmax(q k^T / 8) v per (image, head), S <= 64, d = 64, probabilities recomputed
// All HBM-bound elementwise / row-wise work except the attention backward (small fp32 tiles in shared
// memory, one CTA per (image, head)).
#include "ll_common.cuh"

namespace ll {

// ------------------------------------------------------------------------- transpose (2-byte elements)
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int64_t R, int C, int64_t Rpad)
{
    __shared__ uint16_t tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = r0 + ty + 8 * k;
        const int c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < R && c < C) ? src[r * C + c] : (uint16_t)0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k;
        const int64_t r = r0 + tx;
        if (c < C && r < Rpad) dst[(int64_t)c * Rpad + r] = tile[tx][ty + 8 * k];
    }
}

// ------------------------------------------------------------------------- column sums
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T *__restrict__ src, int64_t R, int C, int64_t rows_per_cta, float *__restrict__ out)
{
    __shared__ float part[4][64];
    const int cx = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int c = blockIdx.x * 64 + cx;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r_end = min(R, r_begin + rows_per_cta);
    float acc = 0.f;
    if (c < C)
        for (int64_t r = r_begin + rg; r < r_end; r += 4) acc += to_float<T>(src[r * C + c]);
    part[rg][cx] = acc;
    __syncthreads();
    if (rg == 0 && c < C) atomicAdd(out + c, part[0][cx] + part[1][cx] + part[2][cx] + part[3][cx]);
}

// ------------------------------------------------------------------------- QuickGELU
template <typename T>
__global__ void __launch_bounds__(256)
quickgelu_forward_kernel(const T *__restrict__ u, T *__restrict__ f, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = to_float<T>(u[i]);
        f[i] = from_float<T>(x / (1.f + __expf(-1.702f * x)));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
quickgelu_backward_kernel(const T *__restrict__ u, const T *__restrict__ g_f, T *__restrict__ g_u, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = to_float<T>(u[i]);
        const float s = 1.f / (1.f + __expf(-1.702f * x));
        // d/dx [x s(1.702 x)] = s + 1.702 x s (1 - s)
        g_u[i] = from_float<T>(to_float<T>(g_f[i]) * (s + 1.702f * x * s * (1.f - s)));
    }
}

// 8 contiguous 16-bit elements <-> 8 floats through one 16-byte access (pointers 16-byte aligned)
template <typename T>
__device__ __forceinline__ void load8(const T *p, float *out)
{
    const uint4 raw = *reinterpret_cast<const uint4 *>(p);
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = to_float<T>(e[k]);
}
template <typename T>
__device__ __forceinline__ void store8(T *p, const float *v)
{
    uint4 raw;
    T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = from_float<T>(v[k]);
    *reinterpret_cast<uint4 *>(p) = raw;
}

// ------------------------------------------------------------------------- LayerNorm backward
// h = (x - mean) * rstd * gamma + beta (clip/model.py LayerNorm: fp32 statistics, eps 1e-5).
//   g_x = g_res + rstd * (g_h*gamma - mean_c(g_h*gamma) - xhat * mean_c(g_h*gamma*xhat))
//   g_gamma += sum_rows g_h * xhat,  g_beta += sum_rows g_h
// One warp per row (grid-stride); a lane owns 8 contiguous columns in each 256-column block.
template <typename T, int NB>
__global__ void __launch_bounds__(256)
layernorm_backward_kernel(const T *__restrict__ x, const T *__restrict__ g_h, const float *__restrict__ gamma,
                          const T *__restrict__ g_res, T *__restrict__ g_x, float *__restrict__ g_gamma,
                          float *__restrict__ g_beta, int64_t rows)
{
    constexpr int W = NB * 256;
    __shared__ float s_gg[W], s_gb[W];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < W; i += 256) s_gg[i] = 0.f, s_gb[i] = 0.f;
    __syncthreads();
    float gam[NB][8], acc_g[NB][8], acc_b[NB][8];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            gam[b][k] = gamma[b * 256 + lane * 8 + k];
            acc_g[b][k] = 0.f;
            acc_b[b][k] = 0.f;
        }
    const int64_t warps_total = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += warps_total) {
        float xv[NB][8], gv[NB][8];
        float sum = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int64_t at = r * W + b * 256 + lane * 8;
            load8<T>(x + at, xv[b]);
            load8<T>(g_h + at, gv[b]);
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += xv[b][k];
        }
        const float mean = warp_sum(sum) * (1.f / W);
        float var = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                xv[b][k] -= mean;
                var += xv[b][k] * xv[b][k];
            }
        const float rstd = rsqrtf(warp_sum(var) * (1.f / W) + 1e-5f);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                xv[b][k] *= rstd;  // xhat
                acc_g[b][k] += gv[b][k] * xv[b][k];
                acc_b[b][k] += gv[b][k];
                gv[b][k] *= gam[b][k];  // d xhat
                m1 += gv[b][k];
                m2 += gv[b][k] * xv[b][k];
            }
        m1 = warp_sum(m1) * (1.f / W);
        m2 = warp_sum(m2) * (1.f / W);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int64_t at = r * W + b * 256 + lane * 8;
            float res[8], g[8];
            if (g_res != nullptr) load8<T>(g_res + at, res);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                g[k] = rstd * (gv[b][k] - m1 - xv[b][k] * m2);
                if (g_res != nullptr) g[k] += res[k];
            }
            store8<T>(g_x + at, g);
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            atomicAdd(&s_gg[b * 256 + lane * 8 + k], acc_g[b][k]);
            atomicAdd(&s_gb[b * 256 + lane * 8 + k], acc_b[b][k]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += 256) {
        atomicAdd(g_gamma + i, s_gg[i]);
        atomicAdd(g_beta + i, s_gb[i]);
    }
}

// ------------------------------------------------------------------------- attention backward
// One CTA (128 threads) per (image, head).  qkv row = [q (H x 64) | k (H x 64) | v (H x 64)].
//   P = softmax(q k^T / 8),  o = P v
//   dV = P^T dO,  dP = dO V^T,  dS = P * (dP - rowsum(P * dP)) / 8,  dQ = dS K,  dK = dS^T Q
constexpr int kAbD = 64;        // head dimension
constexpr int kAbPitch = 66;    // 16-bit elements per staged row (33 words: conflict-free column walks)
constexpr int kAbPPitch = 65;   // floats per probability row

template <typename T>
__global__ void __launch_bounds__(128)
attention_backward_kernel(const T *__restrict__ qkv, const T *__restrict__ g_out, T *__restrict__ g_qkv, int S, int H)
{
    extern __shared__ __align__(16) uint8_t ab_smem[];
    T *sQ = reinterpret_cast<T *>(ab_smem);
    T *sK = sQ + 64 * kAbPitch;
    T *sV = sK + 64 * kAbPitch;
    T *sG = sV + 64 * kAbPitch;
    float *sP = reinterpret_cast<float *>(sG + 64 * kAbPitch);  // [64][65] probabilities
    float *sD = sP + 64 * kAbPPitch;                            // [64][65] dP, then dS
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.x / H, head = blockIdx.x % H;
    const int64_t row0 = (int64_t)img * S;
    const int ld = 3 * H * kAbD, ldo = H * kAbD;
    const float scale = 0.125f;

    // stage q, k, v, dO of this (image, head) as 32-bit words
    for (int i = tid; i < S * 32; i += 128) {
        const int r = i >> 5, w = i & 31;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(qkv + (row0 + r) * ld + head * kAbD);
        reinterpret_cast<uint32_t *>(sQ + r * kAbPitch)[w] = src[w];
        reinterpret_cast<uint32_t *>(sK + r * kAbPitch)[w] = src[w + (H * kAbD) / 2];
        reinterpret_cast<uint32_t *>(sV + r * kAbPitch)[w] = src[w + H * kAbD];
        reinterpret_cast<uint32_t *>(sG + r * kAbPitch)[w] =
            reinterpret_cast<const uint32_t *>(g_out + (row0 + r) * ldo + head * kAbD)[w];
    }
    __syncthreads();
    // scores and dP
    for (int idx = tid; idx < S * S; idx += 128) {
        const int i = idx / S, j = idx % S;
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < kAbD; ++d) {
            s += to_float<T>(sQ[i * kAbPitch + d]) * to_float<T>(sK[j * kAbPitch + d]);
            dp += to_float<T>(sG[i * kAbPitch + d]) * to_float<T>(sV[j * kAbPitch + d]);
        }
        sP[i * kAbPPitch + j] = s * scale;
        sD[i * kAbPPitch + j] = dp;
    }
    __syncthreads();
    // row softmax, delta, dS (a warp per row)
    for (int i = warp; i < S; i += 4) {
        const float s0 = lane < S ? sP[i * kAbPPitch + lane] : -INFINITY;
        const float s1 = lane + 32 < S ? sP[i * kAbPPitch + lane + 32] : -INFINITY;
        const float mx = warp_max(fmaxf(s0, s1));
        const float e0 = lane < S ? __expf(s0 - mx) : 0.f;
        const float e1 = lane + 32 < S ? __expf(s1 - mx) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        const float p0 = e0 * inv, p1 = e1 * inv;
        const float d0 = lane < S ? sD[i * kAbPPitch + lane] : 0.f;
        const float d1 = lane + 32 < S ? sD[i * kAbPPitch + lane + 32] : 0.f;
        const float delta = warp_sum(p0 * d0 + p1 * d1);
        if (lane < S) {
            sP[i * kAbPPitch + lane] = p0;
            sD[i * kAbPPitch + lane] = p0 * (d0 - delta) * scale;
        }
        if (lane + 32 < S) {
            sP[i * kAbPPitch + lane + 32] = p1;
            sD[i * kAbPPitch + lane + 32] = p1 * (d1 - delta) * scale;
        }
    }
    __syncthreads();
    // dQ, dK, dV: a thread owns one (row, column pair)
    for (int idx = tid; idx < S * 32; idx += 128) {
        const int r = idx >> 5, w = idx & 31;
        float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int t = 0; t < S; ++t) {
            const float ds_rt = sD[r * kAbPPitch + t];  // dS[r][t]  (row r as a query)
            const float ds_tr = sD[t * kAbPPitch + r];  // dS[t][r]  (row r as a key)
            const float p_tr = sP[t * kAbPPitch + r];   // P[t][r]
            q0 += ds_rt * to_float<T>(sK[t * kAbPitch + 2 * w]);
            q1 += ds_rt * to_float<T>(sK[t * kAbPitch + 2 * w + 1]);
            k0 += ds_tr * to_float<T>(sQ[t * kAbPitch + 2 * w]);
            k1 += ds_tr * to_float<T>(sQ[t * kAbPitch + 2 * w + 1]);
            v0 += p_tr * to_float<T>(sG[t * kAbPitch + 2 * w]);
            v1 += p_tr * to_float<T>(sG[t * kAbPitch + 2 * w + 1]);
        }
        T *dst = g_qkv + (row0 + r) * ld + head * kAbD + 2 * w;
        dst[0] = from_float<T>(q0);
        dst[1] = from_float<T>(q1);
        dst[H * kAbD] = from_float<T>(k0);
        dst[H * kAbD + 1] = from_float<T>(k1);
        dst[2 * H * kAbD] = from_float<T>(v0);
        dst[2 * H * kAbD + 1] = from_float<T>(v1);
    }
}

constexpr size_t kAbSmem = 4 * 64 * kAbPitch * 2 + 2 * 64 * kAbPPitch * 4;

}  // namespace ll

using namespace ll;

#define LL_DISPATCH_16(dtype, CALL_F16, CALL_BF16, name)                                 \
    switch (dtype) {                                                                     \
        case LL_F16: CALL_F16; break;                                                    \
        case LL_BF16: CALL_BF16; break;                                                  \
        default: return fail(LL_ERR_INVALID, name ": dtype must be LL_F16 or LL_BF16");   \
    }

extern "C" int ll_transpose_pad(const void *src, void *dst, int64_t R, int C, int64_t Rpad, ll_stream_t stream)
{
    LL_REQUIRE(src && dst && R > 0 && C > 0 && Rpad >= R, "transpose_pad: bad arguments");
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((Rpad + 31) / 32));
    transpose_pad_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), R, C, Rpad);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_colsum(const void *src, int64_t R, int C, int dtype, float *out, ll_stream_t stream)
{
    LL_REQUIRE(src && out && R > 0 && C > 0, "colsum: bad arguments");
    const int64_t rows_per_cta = 256;
    dim3 grid((unsigned)((C + 63) / 64), (unsigned)((R + rows_per_cta - 1) / rows_per_cta));
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, (colsum_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half *>(src), R, C, rows_per_cta, out)),
                   (colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(src), R, C, rows_per_cta, out)), "colsum");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_quickgelu_forward(const void *u, void *f, int64_t n, int dtype, ll_stream_t stream)
{
    LL_REQUIRE(u && f && n > 0, "quickgelu_forward: bad arguments");
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, (quickgelu_forward_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half *>(u), static_cast<__half *>(f), n)),
                   (quickgelu_forward_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(u), static_cast<__nv_bfloat16 *>(f), n)), "quickgelu_forward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_quickgelu_backward(const void *u, const void *g_f, void *g_u, int64_t n, int dtype, ll_stream_t stream)
{
    LL_REQUIRE(u && g_f && g_u && n > 0, "quickgelu_backward: bad arguments");
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16);
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, (quickgelu_backward_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half *>(u), static_cast<const __half *>(g_f), static_cast<__half *>(g_u), n)),
                   (quickgelu_backward_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(u), static_cast<const __nv_bfloat16 *>(g_f), static_cast<__nv_bfloat16 *>(g_u), n)), "quickgelu_backward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

template <typename T>
static int launch_ln_bwd(const void *x, const void *g_h, const float *gamma, const void *g_res, void *g_x, float *g_gamma,
                         float *g_beta, int64_t rows, int width, cudaStream_t s)
{
    const unsigned grid = (unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)num_sms() * 2);
    auto X = static_cast<const T *>(x), G = static_cast<const T *>(g_h), R = static_cast<const T *>(g_res);
    auto O = static_cast<T *>(g_x);
    switch (width) {
        case 512: layernorm_backward_kernel<T, 2><<<grid, 256, 0, s>>>(X, G, gamma, R, O, g_gamma, g_beta, rows); break;
        case 768: layernorm_backward_kernel<T, 3><<<grid, 256, 0, s>>>(X, G, gamma, R, O, g_gamma, g_beta, rows); break;
        case 1024: layernorm_backward_kernel<T, 4><<<grid, 256, 0, s>>>(X, G, gamma, R, O, g_gamma, g_beta, rows); break;
        default: return fail(LL_ERR_INVALID, "layernorm_backward: width %d not in {512, 768, 1024}", width);
    }
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_layernorm_backward(const void *x, const void *g_h, const float *gamma, const void *g_res, void *g_x,
                                     float *g_gamma, float *g_beta, int64_t rows, int width, int dtype,
                                     ll_stream_t stream)
{
    LL_REQUIRE(x && g_h && gamma && g_x && g_gamma && g_beta && rows > 0, "layernorm_backward: bad arguments");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_h) | reinterpret_cast<uintptr_t>(g_res) |
                 reinterpret_cast<uintptr_t>(g_x)) & 15) == 0, "layernorm_backward: tensors must be 16-byte aligned");
    if (dtype == LL_F16) return launch_ln_bwd<__half>(x, g_h, gamma, g_res, g_x, g_gamma, g_beta, rows, width, as_stream(stream));
    if (dtype == LL_BF16) return launch_ln_bwd<__nv_bfloat16>(x, g_h, gamma, g_res, g_x, g_gamma, g_beta, rows, width, as_stream(stream));
    return fail(LL_ERR_INVALID, "layernorm_backward: dtype must be LL_F16 or LL_BF16");
}

extern "C" int ll_attention_backward(const void *qkv, const void *g_out, void *g_qkv, int64_t batch, int S, int H,
                                     int dtype, ll_stream_t stream)
{
    LL_REQUIRE(qkv && g_out && g_qkv && batch > 0 && H > 0 && S > 0 && S <= 64, "attention_backward: need 0 < S <= 64");
    LL_REQUIRE(batch * H < (1ll << 31), "attention_backward: batch * heads too large");
    static bool attr_set = false;
    if (!attr_set) {
        LL_CUDA_CHECK(cudaFuncSetAttribute(attention_backward_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAbSmem));
        LL_CUDA_CHECK(cudaFuncSetAttribute(attention_backward_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAbSmem));
        attr_set = true;
    }
    const unsigned grid = (unsigned)(batch * H);
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, (attention_backward_kernel<__half><<<grid, 128, kAbSmem, s>>>(static_cast<const __half *>(qkv), static_cast<const __half *>(g_out), static_cast<__half *>(g_qkv), S, H)),
                   (attention_backward_kernel<__nv_bfloat16><<<grid, 128, kAbSmem, s>>>(static_cast<const __nv_bfloat16 *>(qkv), static_cast<const __nv_bfloat16 *>(g_out), static_cast<__nv_bfloat16 *>(g_qkv), S, H)), "attention_backward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}
