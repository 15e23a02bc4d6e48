// CLIP image preprocessing on the GPU (sm_100a): uint8 images -> bicubic resize (smaller edge to n_px) ->
// centre crop n_px x n_px -> /255 -> (x - mean) / std -> fp16 / fp32 NCHW.
//
// Replaces the CPU DataLoader transform the reference feeds compress_dataset with
// (`clip.clip._transform` returned by clip.load at hub/compressor.py:39, = torchvision Resize(BICUBIC) +
// CenterCrop + ToTensor + Normalize; lossyless/helpers.py:252,260 hold the constants).
// The resize is Pillow's ImagingResample for 8-bit images restated exactly: per output index a window
// [xmin, xmin + n) of integer coefficients (22-bit fixed point, computed on the host in double precision
// with Pillow's formulas), accumulation from 1 << 21, >> 22, clamp to [0, 255], horizontal pass first, then
// vertical on the uint8 intermediate.  Integer work: bit-identical to PIL (tests/test_preprocess_gpu.py).
#include "ll_common.cuh"

namespace ll {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow PRECISION_BITS

__device__ __forceinline__ int clip8(int v)
{
    v >>= kPrecisionBits;
    return min(max(v, 0), 255);
}

// tmp[b, y, xx, c] = resample along x of in[b, y, :, c] for the n_px output columns of the crop window.
// layout 0: in is [B, H, W, 3] (HWC); layout 1: [B, 3, H, W] (CHW).
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t *__restrict__ in, int64_t B, int H, int W, int layout, const int32_t *__restrict__ coef,
                  const int32_t *__restrict__ bounds, int ksize, int n_px, uint8_t *__restrict__ tmp)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = B * H * (int64_t)n_px * 3;
    if (i >= total) return;
    const int c = (int)(i % 3);
    const int xx = (int)((i / 3) % n_px);
    const int y = (int)((i / (3 * (int64_t)n_px)) % H);
    const int64_t b = i / (3 * (int64_t)n_px * H);
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int32_t *k = coef + (int64_t)xx * ksize;
    int ss = 1 << (kPrecisionBits - 1);
    if (layout == 0) {
        const uint8_t *row = in + ((b * H + y) * (int64_t)W + xmin) * 3 + c;
        for (int x = 0; x < n; ++x) ss += (int)row[3 * x] * k[x];
    } else {
        const uint8_t *row = in + ((b * 3 + c) * (int64_t)H + y) * W + xmin;
        for (int x = 0; x < n; ++x) ss += (int)row[x] * k[x];
    }
    tmp[i] = (uint8_t)clip8(ss);
}

// out[b, c, yy, xx] = normalise(resample along y of tmp[b, :, xx, c]) for the n_px output rows of the crop window.
template <typename TOut>
__global__ void __launch_bounds__(256)
resample_v_normalize_kernel(const uint8_t *__restrict__ tmp, int64_t B, int H, const int32_t *__restrict__ coef,
                            const int32_t *__restrict__ bounds, int ksize, int n_px, float m0, float m1, float m2, float s0,
                            float s1, float s2, TOut *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = B * 3 * (int64_t)n_px * n_px;
    if (i >= total) return;
    const int xx = (int)(i % n_px);
    const int yy = (int)((i / n_px) % n_px);
    const int c = (int)((i / ((int64_t)n_px * n_px)) % 3);
    const int64_t b = i / (3 * (int64_t)n_px * n_px);
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int32_t *k = coef + (int64_t)yy * ksize;
    const uint8_t *col = tmp + ((b * H + ymin) * (int64_t)n_px + xx) * 3 + c;
    int ss = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < n; ++y) ss += (int)col[(int64_t)y * n_px * 3] * k[y];
    const int px = clip8(ss);
    // ToTensor: uint8 -> float32 / 255 ; Normalize: (x - mean) / std, each a single IEEE float32 rounding
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float std = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px, 255.f), mean), std);
    out[i] = from_float<TOut>(v);
}

// Fused variant for images small enough to live in shared memory (STL10 96x96, CIFAR, MNIST ...: the input,
// the horizontally resampled intermediate and the coefficient tables together fit 227 KB): one CTA per
// image, both passes out of shared memory, normalisation through a 3 x 256 look-up table built with the
// same IEEE operations as above (so the results are the same bits), 2 output pixels per thread and store.
// The two-kernel path above wrote and re-read the intermediate through L2 with one thread per output byte.
template <typename TOut>
struct Pair;
template <>
struct Pair<float> {
    using type = float2;
    static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};
template <>
struct Pair<__half> {
    using type = __half2;
    static __device__ __forceinline__ __half2 make(__half a, __half b) { return __halves2half2(a, b); }
};
template <>
struct Pair<__nv_bfloat16> {
    using type = __nv_bfloat162;
    static __device__ __forceinline__ __nv_bfloat162 make(__nv_bfloat16 a, __nv_bfloat16 b) { return __halves2bfloat162(a, b); }
};

constexpr int kFusedThreads = 512;

struct FusedLayout {
    int in_off, tmp_off, ch_off, cv_off, bh_off, bv_off, lut_off, total;
};
static __host__ __device__ inline FusedLayout fused_layout(int H, int W, int n_px, int ksize_h, int ksize_v, int out_elem)
{
    auto up = [](int v) { return (v + 15) & ~15; };
    FusedLayout l;
    int off = 0;
    l.in_off = off, off += up(H * W * 3 + 16);  // + 16: room to keep the global 16-byte phase of the image
    l.tmp_off = off, off += up(H * n_px * 3);
    l.ch_off = off, off += up(n_px * ksize_h * 4);
    l.cv_off = off, off += up(n_px * ksize_v * 4);
    l.bh_off = off, off += up(n_px * 8);
    l.bv_off = off, off += up(n_px * 8);
    l.lut_off = off, off += up(3 * 256 * out_elem);
    l.total = off;
    return l;
}

template <typename TOut>
__global__ void __launch_bounds__(kFusedThreads)
clip_preprocess_fused_kernel(const uint8_t *__restrict__ img, int H, int W, int layout, const int32_t *__restrict__ coef_h,
                             const int32_t *__restrict__ bounds_h, int ksize_h, const int32_t *__restrict__ coef_v,
                             const int32_t *__restrict__ bounds_v, int ksize_v, int n_px, float m0, float m1, float m2, float s0,
                             float s1, float s2, TOut *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const FusedLayout L = fused_layout(H, W, n_px, ksize_h, ksize_v, (int)sizeof(TOut));
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    const int nbytes = H * W * 3;
    const uint8_t *src = img + b * nbytes;
    // the image lands at the same 16-byte phase it has in global memory, so the body copies as uint4
    const int phase = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    uint8_t *s_in = smem + L.in_off + phase;
    uint8_t *s_tmp = smem + L.tmp_off;
    int32_t *s_ch = reinterpret_cast<int32_t *>(smem + L.ch_off), *s_cv = reinterpret_cast<int32_t *>(smem + L.cv_off);
    int2 *s_bh = reinterpret_cast<int2 *>(smem + L.bh_off), *s_bv = reinterpret_cast<int2 *>(smem + L.bv_off);
    TOut *s_lut = reinterpret_cast<TOut *>(smem + L.lut_off);
    {
        const int head = min(nbytes, (16 - phase) & 15);
        for (int i = tid; i < head; i += kFusedThreads) s_in[i] = src[i];
        const int nvec = (nbytes - head) >> 4;
        const uint4 *gv = reinterpret_cast<const uint4 *>(src + head);
        uint4 *sv = reinterpret_cast<uint4 *>(s_in + head);
        for (int i = tid; i < nvec; i += kFusedThreads) sv[i] = __ldg(gv + i);
        for (int i = head + (nvec << 4) + tid; i < nbytes; i += kFusedThreads) s_in[i] = src[i];
    }
    for (int i = tid; i < n_px * ksize_h; i += kFusedThreads) s_ch[i] = coef_h[i];
    for (int i = tid; i < n_px * ksize_v; i += kFusedThreads) s_cv[i] = coef_v[i];
    for (int i = tid; i < n_px; i += kFusedThreads) {
        s_bh[i] = make_int2(bounds_h[2 * i], bounds_h[2 * i + 1]);
        s_bv[i] = make_int2(bounds_v[2 * i], bounds_v[2 * i + 1]);
    }
    for (int i = tid; i < 3 * 256; i += kFusedThreads) {
        const int c = i >> 8;
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), std = c == 0 ? s0 : (c == 1 ? s1 : s2);
        // ToTensor: uint8 -> float32 / 255 ; Normalize: (x - mean) / std, each a single IEEE float32 rounding
        s_lut[i] = from_float<TOut>(__fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.f), mean), std));
    }
    __syncthreads();
    // horizontal pass: s_tmp[y, xx, c]
    const int stride = layout == 0 ? 3 : 1;
    for (int i = tid; i < H * n_px * 3; i += kFusedThreads) {
        const int c = i % 3, xx = (i / 3) % n_px, y = i / (3 * n_px);
        const int2 bd = s_bh[xx];
        const int32_t *k = s_ch + xx * ksize_h;
        const uint8_t *p = layout == 0 ? s_in + (y * W + bd.x) * 3 + c : s_in + (c * H + y) * W + bd.x;
        int ss = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < bd.y; ++x) ss += (int)p[x * stride] * k[x];
        s_tmp[i] = (uint8_t)clip8(ss);
    }
    __syncthreads();
    // vertical pass + normalise: two neighbouring output pixels per thread
    const int half = n_px >> 1;
    typename Pair<TOut>::type *dst = reinterpret_cast<typename Pair<TOut>::type *>(out + b * 3 * (int64_t)n_px * n_px);
    for (int i = tid; i < 3 * n_px * half; i += kFusedThreads) {
        const int xp = i % half, yy = (i / half) % n_px, c = i / (half * n_px);
        const int2 bd = s_bv[yy];
        const int32_t *k = s_cv + yy * ksize_v;
        const uint8_t *col = s_tmp + (bd.x * n_px + 2 * xp) * 3 + c;
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0;
        for (int y = 0; y < bd.y; ++y) {
            const int kk = k[y];
            a0 += (int)col[y * n_px * 3] * kk;
            a1 += (int)col[y * n_px * 3 + 3] * kk;
        }
        dst[i] = Pair<TOut>::make(s_lut[(c << 8) + clip8(a0)], s_lut[(c << 8) + clip8(a1)]);
    }
}

// Identity-size variant (H == W == n_px: torchvision's Resize returns the image unchanged and the centre crop is the
// whole image): only ToTensor + Normalize + the cast remain -- uint8 HWC (or CHW) in, 16-/32-bit CHW out.  This is the
// CLIP-sized uint8 feed of compress_dataset (150 KB per image over the host link instead of 301 KB of fp16).  One CTA
// per 8 image rows: 16-byte loads into shared memory, normalisation through the same 3 x 256 IEEE look-up table as the
// fused kernel above (same bits), 16-byte stores of 8 (16-bit) or 4 (fp32) neighbouring pixels of one channel plane.
constexpr int kIdRows = 8;
constexpr int kIdThreads = 256;

template <typename TOut>
__global__ void __launch_bounds__(kIdThreads)
clip_normalize_u8_kernel(const uint8_t *__restrict__ img, int n_px, int layout, float m0, float m1, float m2, float s0, float s1,
                         float s2, TOut *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kPerVec = 16 / (int)sizeof(TOut);
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    const int y0 = blockIdx.x * kIdRows;
    const int rows = min(kIdRows, n_px - y0);
    TOut *s_lut = reinterpret_cast<TOut *>(smem);
    uint8_t *s_in = smem + 3 * 256 * sizeof(TOut);  // [3 planes or 1 HWC slab][rows * n_px (* 3)], 16-byte aligned pieces
    for (int i = tid; i < 3 * 256; i += kIdThreads) {
        const int c = i >> 8;
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), std = c == 0 ? s0 : (c == 1 ? s1 : s2);
        s_lut[i] = from_float<TOut>(__fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.f), mean), std));
    }
    const int plane = n_px * n_px;
    if (layout == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(img + (b * plane + (int64_t)y0 * n_px) * 3);
        uint4 *dst = reinterpret_cast<uint4 *>(s_in);
        for (int i = tid; i < (rows * n_px * 3) >> 4; i += kIdThreads) dst[i] = __ldg(src + i);
    } else {
        for (int c = 0; c < 3; ++c) {
            const uint4 *src = reinterpret_cast<const uint4 *>(img + (b * 3 + c) * (int64_t)plane + (int64_t)y0 * n_px);
            uint4 *dst = reinterpret_cast<uint4 *>(s_in + c * kIdRows * n_px);
            for (int i = tid; i < (rows * n_px) >> 4; i += kIdThreads) dst[i] = __ldg(src + i);
        }
    }
    __syncthreads();
    const int vec_per_row = n_px / kPerVec;
    for (int i = tid; i < 3 * rows * vec_per_row; i += kIdThreads) {
        const int xv = i % vec_per_row, r = (i / vec_per_row) % rows, c = i / (vec_per_row * rows);
        alignas(16) TOut v[kPerVec];
#pragma unroll
        for (int k = 0; k < kPerVec; ++k) {
            const int x = xv * kPerVec + k;
            const int px = layout == 0 ? s_in[(r * n_px + x) * 3 + c] : s_in[c * kIdRows * n_px + r * n_px + x];
            v[k] = s_lut[(c << 8) + px];
        }
        *reinterpret_cast<uint4 *>(out + ((b * 3 + c) * (int64_t)plane + (int64_t)(y0 + r) * n_px) + xv * kPerVec) =
            *reinterpret_cast<const uint4 *>(v);
    }
}

template <typename TOut>
static int launch_identity(const uint8_t *img, int64_t B, int n_px, int layout, const float *mean3, const float *std3, TOut *out,
                           cudaStream_t s)
{
    const int smem = 3 * 256 * (int)sizeof(TOut) + 3 * kIdRows * n_px;
    dim3 grid((unsigned)((n_px + kIdRows - 1) / kIdRows), (unsigned)B);
    clip_normalize_u8_kernel<TOut><<<grid, kIdThreads, smem, s>>>(img, n_px, layout, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                                                std3[2], out);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

template <typename TOut>
static int launch_fused(const uint8_t *img, int64_t B, int H, int W, int layout, const int32_t *coef_h, const int32_t *bounds_h,
                        int ksize_h, const int32_t *coef_v, const int32_t *bounds_v, int ksize_v, const float *mean3,
                        const float *std3, int n_px, TOut *out, int smem, cudaStream_t s)
{
    auto kern = clip_preprocess_fused_kernel<TOut>;
    LL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)B, kFusedThreads, smem, s>>>(img, H, W, layout, coef_h, bounds_h, ksize_h, coef_v, bounds_v, ksize_v, n_px,
                                                  mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

}  // namespace ll

using namespace ll;

static int g_preprocess_impl = 0;  // 0: fused shared-memory kernel when the image fits, 1: always the two-pass kernels
extern "C" int ll_clip_preprocess_set_impl(int impl)
{
    LL_REQUIRE(impl == 0 || impl == 1, "clip_preprocess_set_impl: impl must be 0 or 1");
    g_preprocess_impl = impl;
    return LL_OK;
}

extern "C" int ll_clip_preprocess(const uint8_t *img, int64_t B, int H, int W, int layout, const int32_t *coef_h,
                                  const int32_t *bounds_h, int ksize_h, const int32_t *coef_v, const int32_t *bounds_v,
                                  int ksize_v, const float *mean3, const float *std3, int n_px, uint8_t *tmp, void *out,
                                  int out_dtype, ll_stream_t stream)
{
    LL_REQUIRE(img && coef_h && bounds_h && coef_v && bounds_v && mean3 && std3 && tmp && out, "clip_preprocess: null pointer");
    LL_REQUIRE(B >= 0 && H > 0 && W > 0 && n_px > 0 && ksize_h > 0 && ksize_v > 0 && (layout == 0 || layout == 1), "clip_preprocess: bad arguments");
    if (B == 0) return LL_OK;
    cudaStream_t s = as_stream(stream);
    // no resize needed (Resize(n_px) leaves an n_px x n_px image unchanged; bicubic coefficients at scale 1 are the
    // identity, so the general path gives the same bytes): normalise only
    if (g_preprocess_impl != 1 && H == n_px && W == n_px && n_px % 16 == 0 && n_px * 3 * kIdRows <= 40 * 1024 && B <= 65535 &&
        (reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        switch (out_dtype) {
            case LL_F32: return launch_identity<float>(img, B, n_px, layout, mean3, std3, (float *)out, s);
            case LL_F16: return launch_identity<__half>(img, B, n_px, layout, mean3, std3, (__half *)out, s);
            case LL_BF16: return launch_identity<__nv_bfloat16>(img, B, n_px, layout, mean3, std3, (__nv_bfloat16 *)out, s);
            default: return fail(LL_ERR_INVALID, "clip_preprocess: unknown output dtype %d", out_dtype);
        }
    }
    // small images: both passes out of shared memory, one CTA per image
    const int out_elem = out_dtype == LL_F32 ? 4 : 2;
    const FusedLayout fl = fused_layout(H, W, n_px, ksize_h, ksize_v, out_elem);
    if (g_preprocess_impl != 1 && (int64_t)H * W * 3 + (int64_t)H * n_px * 3 < 200000 && fl.total <= 227 * 1024 && n_px % 2 == 0 &&
        B <= 0x7fffffff && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
        switch (out_dtype) {
            case LL_F32: return launch_fused<float>(img, B, H, W, layout, coef_h, bounds_h, ksize_h, coef_v, bounds_v, ksize_v, mean3, std3, n_px, (float *)out, fl.total, s);
            case LL_F16: return launch_fused<__half>(img, B, H, W, layout, coef_h, bounds_h, ksize_h, coef_v, bounds_v, ksize_v, mean3, std3, n_px, (__half *)out, fl.total, s);
            case LL_BF16: return launch_fused<__nv_bfloat16>(img, B, H, W, layout, coef_h, bounds_h, ksize_h, coef_v, bounds_v, ksize_v, mean3, std3, n_px, (__nv_bfloat16 *)out, fl.total, s);
            default: return fail(LL_ERR_INVALID, "clip_preprocess: unknown output dtype %d", out_dtype);
        }
    }
    const int64_t t1 = B * H * (int64_t)n_px * 3;
    resample_h_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, s>>>(img, B, H, W, layout, coef_h, bounds_h, ksize_h, n_px, tmp);
    LL_LAUNCH_CHECK();
    const int64_t t2 = B * 3 * (int64_t)n_px * n_px;
    const unsigned grid = (unsigned)((t2 + 255) / 256);
    switch (out_dtype) {
        case LL_F32: resample_v_normalize_kernel<float><<<grid, 256, 0, s>>>(tmp, B, H, coef_v, bounds_v, ksize_v, n_px, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], (float *)out); break;
        case LL_F16: resample_v_normalize_kernel<__half><<<grid, 256, 0, s>>>(tmp, B, H, coef_v, bounds_v, ksize_v, n_px, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], (__half *)out); break;
        case LL_BF16: resample_v_normalize_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(tmp, B, H, coef_v, bounds_v, ksize_v, n_px, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], (__nv_bfloat16 *)out); break;
        default: return fail(LL_ERR_INVALID, "clip_preprocess: unknown output dtype %d", out_dtype);
    }
    LL_LAUNCH_CHECK();
    return LL_OK;
}
