// Flat-buffer optimiser kernels for the fine-tuning path (sm_100a): everything the training step does to the 88 M
// parameters of an unfrozen CLIP tower besides the GEMMs, as three streaming kernels over ONE flat buffer per dtype
// instead of ~10 launches per parameter tensor (ATen casts, accumulations, foreach norms, fused-AdamW chunks: 858 + 36 +
// 80 + 51 launches = 15 % of the step in profiles/r01_ncu_launches_train_unfrozen.txt).
//
// Replaces, for the encoder's parameters: the autocast casts of lossyless/architectures.py:351-376 under
// `precision: 16` (config/main.yaml:124), `gradient_clip_val` (config/main.yaml:113 -> torch clip_grad_norm_) and
// torch.optim.AdamW (config/featurizer/bottleneck_clip_lossyZ.yaml:37-47) as driven by
// lossyless/learnable_compressors.py:277-303.
//
//   ll_sumsq        sum of squares of a flat fp32 / 16-bit buffer -> one float (two stages, fixed order: deterministic)
//   ll_adamw_flat   p, m, v <- AdamW(p, g * scale), optionally also the 16-bit compute copy of p; `scale` is read from
//                   device memory (clip coefficient / world size, computed on the device: no host synchronisation)
//   ll_cast_flat    fp32 -> 16-bit copy (the compute copies, when parameters were changed by something else)
// All HBM-bound: 16-byte accesses, grid = a multiple of the SM count.
#include "ll_common.cuh"

namespace ll {

constexpr int kOptThreads = 256;

template <typename T>
struct Vec8;  // eight consecutive elements <-> floats
template <>
struct Vec8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&f)[8])
    {
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&f)[8])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
};
template <>
struct Vec8<__half> {
    static __device__ __forceinline__ void load(const __half *p, float (&f)[8])
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(h[i]);
            f[2 * i] = t.x, f[2 * i + 1] = t.y;
        }
    }
    static __device__ __forceinline__ void store(__half *p, const float (&f)[8])
    {
        uint4 v;
        __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = v;
    }
};
template <>
struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[8])
    {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __bfloat1622float2(h[i]);
            f[2 * i] = t.x, f[2 * i + 1] = t.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[8])
    {
        uint4 v;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        *reinterpret_cast<uint4 *>(p) = v;
    }
};

__device__ __forceinline__ float block_sum(float v, float *sm)
{
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.f;
        v = warp_sum(v);
    }
    return v;  // valid in warp 0
}

// stage 1: partial[b] = sum over the block's grid-stride share; n8 = number of 8-element vectors
template <typename T>
__global__ void __launch_bounds__(kOptThreads)
sumsq_partial_kernel(const T *__restrict__ x, int64_t n8, float *__restrict__ partial)
{
    __shared__ float sm[32];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kOptThreads) {
        float f[8];
        Vec8<T>::load(x + i * 8, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(f[e], f[e], acc);
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// stage 2: out[0] (+)= sum of the partials in index order
__global__ void __launch_bounds__(kOptThreads)
sumsq_final_kernel(const float *__restrict__ partial, int n, float *__restrict__ out, int accumulate)
{
    __shared__ float sm[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += kOptThreads) acc += partial[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + acc : acc;
}

// torch.optim.AdamW (decoupled weight decay, bias-corrected moments), float32 state:
//   g  = grad * scale[0]
//   p *= 1 - lr * wd ;  m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
template <typename TG, typename TW>
__global__ void __launch_bounds__(kOptThreads)
adamw_flat_kernel(float *__restrict__ p, const TG *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                  TW *__restrict__ w16, int64_t n8, float lr, float b1, float b2, float eps, float wd, float bc1, float rsqrt_bc2,
                  const float *__restrict__ scale)
{
    const float s = scale != nullptr ? scale[0] : 1.f;
    const float decay = 1.f - lr * wd, step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kOptThreads) {
        float pf[8], gf[8], mf[8], vf[8];
        Vec8<float>::load(p + i * 8, pf);
        Vec8<TG>::load(g + i * 8, gf);
        Vec8<float>::load(m + i * 8, mf);
        Vec8<float>::load(v + i * 8, vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float ge = gf[e] * s;
            mf[e] = fmaf(b1, mf[e], (1.f - b1) * ge);
            vf[e] = fmaf(b2, vf[e], (1.f - b2) * ge * ge);
            const float denom = fmaf(sqrtf(vf[e]), rsqrt_bc2, eps);
            pf[e] = fmaf(-step_size, __fdividef(mf[e], denom), pf[e] * decay);
        }
        Vec8<float>::store(p + i * 8, pf);
        Vec8<float>::store(m + i * 8, mf);
        Vec8<float>::store(v + i * 8, vf);
        if (w16 != nullptr) Vec8<TW>::store(w16 + i * 8, pf);
    }
}

template <typename TW>
__global__ void __launch_bounds__(kOptThreads)
cast_flat_kernel(const float *__restrict__ src, TW *__restrict__ dst, int64_t n8)
{
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n8; i += (int64_t)gridDim.x * kOptThreads) {
        float f[8];
        Vec8<float>::load(src + i * 8, f);
        Vec8<TW>::store(dst + i * 8, f);
    }
}

static inline int stream_grid(int64_t n8)
{
    const int64_t want = (n8 + kOptThreads - 1) / kOptThreads;
    const int64_t cap = 8LL * num_sms();  // 8 CTAs of 256 threads per SM: one full wave
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace ll

using namespace ll;

extern "C" int ll_sumsq_scratch_floats(void) { return 8 * 148 * 2; }

extern "C" int ll_sumsq(const void *x, int64_t n, int dtype, float *scratch, float *out, int accumulate, ll_stream_t stream)
{
    LL_REQUIRE(x && scratch && out && n >= 0 && n % 8 == 0, "sumsq: needs n %% 8 == 0 and non-null pointers (n=%lld)", (long long)n);
    LL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: x must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int64_t n8 = n / 8;
    int grid = stream_grid(n8);
    if (grid > ll_sumsq_scratch_floats()) grid = ll_sumsq_scratch_floats();
    switch (dtype) {
        case LL_F32: sumsq_partial_kernel<float><<<grid, kOptThreads, 0, s>>>((const float *)x, n8, scratch); break;
        case LL_F16: sumsq_partial_kernel<__half><<<grid, kOptThreads, 0, s>>>((const __half *)x, n8, scratch); break;
        case LL_BF16: sumsq_partial_kernel<__nv_bfloat16><<<grid, kOptThreads, 0, s>>>((const __nv_bfloat16 *)x, n8, scratch); break;
        default: return fail(LL_ERR_INVALID, "sumsq: unknown dtype %d", dtype);
    }
    LL_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, kOptThreads, 0, s>>>(scratch, grid, out, accumulate);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_adamw_flat(float *p, const void *g, int g_dtype, float *m, float *v, void *w16, int w16_dtype, int64_t n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                             const float *grad_scale, ll_stream_t stream)
{
    LL_REQUIRE(p && g && m && v && n >= 0 && n % 8 == 0 && step >= 1, "adamw_flat: bad arguments (n=%lld step=%lld)", (long long)n, (long long)step);
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(w16)) & 15) == 0, "adamw_flat: buffers must be 16-byte aligned");
    if (n == 0) return LL_OK;
    cudaStream_t s = as_stream(stream);
    const int64_t n8 = n / 8;
    const int grid = stream_grid(n8);
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float rsqrt_bc2 = 1.f / sqrtf(1.f - powf(beta2, (float)step));
#define LL_ADAMW(TG, TW) adamw_flat_kernel<TG, TW><<<grid, kOptThreads, 0, s>>>(p, (const TG *)g, m, v, (TW *)w16, n8, lr, beta1, beta2, eps, weight_decay, bc1, rsqrt_bc2, grad_scale)
    const bool w_half = w16 != nullptr && w16_dtype == LL_F16;
    LL_REQUIRE(w16 == nullptr || w16_dtype == LL_F16 || w16_dtype == LL_BF16, "adamw_flat: the compute copy must be fp16 or bf16");
    switch (g_dtype) {
        case LL_F32: if (w_half) LL_ADAMW(float, __half); else LL_ADAMW(float, __nv_bfloat16); break;
        case LL_F16: if (w_half) LL_ADAMW(__half, __half); else LL_ADAMW(__half, __nv_bfloat16); break;
        case LL_BF16: if (w_half) LL_ADAMW(__nv_bfloat16, __half); else LL_ADAMW(__nv_bfloat16, __nv_bfloat16); break;
        default: return fail(LL_ERR_INVALID, "adamw_flat: unknown gradient dtype %d", g_dtype);
    }
#undef LL_ADAMW
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_cast_flat(const float *src, void *dst, int dst_dtype, int64_t n, ll_stream_t stream)
{
    LL_REQUIRE(src && dst && n >= 0 && n % 8 == 0, "cast_flat: needs n %% 8 == 0 and non-null pointers");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "cast_flat: buffers must be 16-byte aligned");
    if (n == 0) return LL_OK;
    cudaStream_t s = as_stream(stream);
    const int64_t n8 = n / 8;
    const int grid = stream_grid(n8);
    switch (dst_dtype) {
        case LL_F16: cast_flat_kernel<__half><<<grid, kOptThreads, 0, s>>>(src, (__half *)dst, n8); break;
        case LL_BF16: cast_flat_kernel<__nv_bfloat16><<<grid, kOptThreads, 0, s>>>(src, (__nv_bfloat16 *)dst, n8); break;
        default: return fail(LL_ERR_INVALID, "cast_flat: destination must be fp16 or bf16");
    }
    LL_LAUNCH_CHECK();
    return LL_OK;
}
