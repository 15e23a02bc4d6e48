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

}  // namespace ll

using namespace ll;

extern "C" int ll_clip_preprocess(const uint8_t *img, int64_t B, int H, int W, int layout, const int32_t *coef_h,
                                  const int32_t *bounds_h, int ksize_h, const int32_t *coef_v, const int32_t *bounds_v,
                                  int ksize_v, const float *mean3, const float *std3, int n_px, uint8_t *tmp, void *out,
                                  int out_dtype, ll_stream_t stream)
{
    LL_REQUIRE(img && coef_h && bounds_h && coef_v && bounds_v && mean3 && std3 && tmp && out, "clip_preprocess: null pointer");
    LL_REQUIRE(B >= 0 && H > 0 && W > 0 && n_px > 0 && ksize_h > 0 && ksize_v > 0 && (layout == 0 || layout == 1), "clip_preprocess: bad arguments");
    if (B == 0) return LL_OK;
    cudaStream_t s = as_stream(stream);
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
