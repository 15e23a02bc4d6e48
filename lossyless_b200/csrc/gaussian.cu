// Gaussian-conditional likelihood, forward and backward, fused (sm_100a).
//
// Replaces compressai==1.1.5 GaussianConditional.forward (quantize "noise" / "dequantize", _likelihood =
// Phi((1/2 - |y - mu|) / s) - Phi((-1/2 - |y - mu|) / s) with the scale's LowerBound(0.11), the likelihood's
// LowerBound(1e-9)) plus the per-row -log q sum of the reference's scale-hyperprior rate
// (lossyless/rates.py:625-650, HRateHyperprior.forward_help) and their autograd: ~25 elementwise ATen launches forward,
// about twice that backward, in ONE launch each.
//
// Layout: x / scales / means / noise [rows, C] float32 row-major; one warp per row, a lane owns 4 consecutive columns
// (128-bit accesses), row reductions by warp shuffle.  HBM-bound: 16-20 B read + 8 B written per element forward.
// Phi(t) = erfc(-t / sqrt 2) / 2 with libdevice erfcf (the complementary form keeps the tails' relative precision,
// as compressai's _standardized_cumulative does); its derivative is the normal density.
#include "ll_common.cuh"

namespace ll {

constexpr int kGcThreads = 256;

__device__ __forceinline__ float std_cdf(float t) { return 0.5f * erfcf(-0.70710678118654752f * t); }
__device__ __forceinline__ float std_pdf(float t) { return 0.3989422804014327f * __expf(-0.5f * t * t); }

struct GcElem {
    float out, v, sgn, s, upper, lower, lik, q;
    bool s_pass;  // the scale sat at or above its bound
};

__device__ __forceinline__ GcElem gc_eval(float x, float scale, float mean, const float *noise_p, float scale_bound, float lik_bound)
{
    GcElem e;
    e.out = noise_p != nullptr ? x + *noise_p : __fadd_rn(rintf(__fsub_rn(x, mean)), mean);  // quantize "noise" / "dequantize"
    const float d = e.out - mean;
    e.v = fabsf(d);
    e.sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    e.s_pass = scale >= scale_bound;
    e.s = fmaxf(scale, scale_bound);
    e.upper = std_cdf((0.5f - e.v) / e.s);
    e.lower = std_cdf((-0.5f - e.v) / e.s);
    e.lik = e.upper - e.lower;
    e.q = fmaxf(e.lik, lik_bound);
    return e;
}

__global__ void __launch_bounds__(kGcThreads)
gc_forward_kernel(const float *__restrict__ x, const float *__restrict__ scales, const float *__restrict__ means,
                  const float *__restrict__ noise, int64_t rows, int C, float scale_bound, float lik_bound,
                  float *__restrict__ out, float *__restrict__ q, float *__restrict__ rates)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kGcThreads / 32);
    for (int64_t r = (int64_t)blockIdx.x * (kGcThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
        float nl = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const int64_t at = r * C + c;
            const float4 xv = *reinterpret_cast<const float4 *>(x + at);
            const float4 sv = *reinterpret_cast<const float4 *>(scales + at);
            const float4 mv = means != nullptr ? *reinterpret_cast<const float4 *>(means + at) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (noise != nullptr) nv = *reinterpret_cast<const float4 *>(noise + at);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w},
                        ns[4] = {nv.x, nv.y, nv.z, nv.w};
            float o[4], qq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const GcElem e = gc_eval(xs[k], ss[k], ms[k], noise != nullptr ? &ns[k] : nullptr, scale_bound, lik_bound);
                o[k] = e.out, qq[k] = e.q;
                nl -= logf(e.q);
            }
            *reinterpret_cast<float4 *>(out + at) = make_float4(o[0], o[1], o[2], o[3]);
            if (q != nullptr) *reinterpret_cast<float4 *>(q + at) = make_float4(qq[0], qq[1], qq[2], qq[3]);
        }
        if (rates != nullptr) {
            nl = warp_sum(nl);
            if (lane == 0) rates[r] = nl;
        }
    }
}

// Incoming: g_out (dL/d out), g_q (dL/d q), g_rates (dL/d rates[r], rates = -sum log q); any may be null.
//   g_lik = (g_q - g_rates / q) through LowerBound (passes where lik >= bound or the gradient is negative)
//   d lik / d v = (pdf(b) - pdf(a)) / s,   d lik / d s = (b pdf(b) - a pdf(a)) / s,   a = (1/2 - v) / s, b = (-1/2 - v) / s
//   v = |out - mean|;  training: out = x + noise;  eval: out = round(x - mean) + mean (v does not move with x or mean)
__global__ void __launch_bounds__(kGcThreads)
gc_backward_kernel(const float *__restrict__ x, const float *__restrict__ scales, const float *__restrict__ means,
                   const float *__restrict__ noise, int64_t rows, int C, float scale_bound, float lik_bound,
                   const float *__restrict__ g_out, const float *__restrict__ g_q, const float *__restrict__ g_rates,
                   float *__restrict__ gx, float *__restrict__ gscales, float *__restrict__ gmeans)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (kGcThreads / 32);
    for (int64_t r = (int64_t)blockIdx.x * (kGcThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
        const float gr = g_rates != nullptr ? g_rates[r] : 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const int64_t at = r * C + c;
            const float4 xv = *reinterpret_cast<const float4 *>(x + at);
            const float4 sv = *reinterpret_cast<const float4 *>(scales + at);
            const float4 mv = means != nullptr ? *reinterpret_cast<const float4 *>(means + at) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 nv = make_float4(0.f, 0.f, 0.f, 0.f), gov = nv, gqv = nv;
            if (noise != nullptr) nv = *reinterpret_cast<const float4 *>(noise + at);
            if (g_out != nullptr) gov = *reinterpret_cast<const float4 *>(g_out + at);
            if (g_q != nullptr) gqv = *reinterpret_cast<const float4 *>(g_q + at);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w},
                        ns[4] = {nv.x, nv.y, nv.z, nv.w}, go[4] = {gov.x, gov.y, gov.z, gov.w}, gq[4] = {gqv.x, gqv.y, gqv.z, gqv.w};
            float dx[4], ds[4], dm[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const GcElem e = gc_eval(xs[k], ss[k], ms[k], noise != nullptr ? &ns[k] : nullptr, scale_bound, lik_bound);
                float g_lik = gq[k] - gr / e.q;
                if (!(e.lik >= lik_bound || g_lik < 0.f)) g_lik = 0.f;
                const float a = (0.5f - e.v) / e.s, b = (-0.5f - e.v) / e.s;
                const float pa = std_pdf(a), pb = std_pdf(b);
                const float g_v = g_lik * (pb - pa) / e.s;
                float g_s = g_lik * (b * pb - a * pa) / e.s;
                if (!(e.s_pass || g_s < 0.f)) g_s = 0.f;  // LowerBound on the scale
                ds[k] = g_s;
                if (noise != nullptr) {
                    dx[k] = go[k] + g_v * e.sgn;   // out = x + noise
                    dm[k] = -g_v * e.sgn;
                } else {
                    dx[k] = 0.f;                   // round(): no gradient to x; out moves 1:1 with the mean, v does not
                    dm[k] = go[k];
                }
            }
            *reinterpret_cast<float4 *>(gx + at) = make_float4(dx[0], dx[1], dx[2], dx[3]);
            *reinterpret_cast<float4 *>(gscales + at) = make_float4(ds[0], ds[1], ds[2], ds[3]);
            if (gmeans != nullptr) *reinterpret_cast<float4 *>(gmeans + at) = make_float4(dm[0], dm[1], dm[2], dm[3]);
        }
    }
}

static inline unsigned gc_grid(int64_t rows)
{
    const int64_t want = (rows + kGcThreads / 32 - 1) / (kGcThreads / 32);
    const int64_t cap = 8LL * num_sms();
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace ll

using namespace ll;

extern "C" int ll_gc_forward(const float *x, const float *scales, const float *means, const float *noise, int64_t rows, int C,
                             float scale_bound, float likelihood_bound, float *out, float *q, float *rates, ll_stream_t stream)
{
    LL_REQUIRE(x && scales && out && rows >= 0 && C > 0 && C % 4 == 0, "gc_forward: needs C %% 4 == 0 and non-null x / scales / out");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(means) |
                 reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(q)) & 15) == 0,
               "gc_forward: tensors must be 16-byte aligned");
    if (rows == 0) return LL_OK;
    gc_forward_kernel<<<gc_grid(rows), kGcThreads, 0, as_stream(stream)>>>(x, scales, means, noise, rows, C, scale_bound, likelihood_bound, out, q, rates);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_gc_backward(const float *x, const float *scales, const float *means, const float *noise, int64_t rows, int C,
                              float scale_bound, float likelihood_bound, const float *g_out, const float *g_q, const float *g_rates,
                              float *gx, float *gscales, float *gmeans, ll_stream_t stream)
{
    LL_REQUIRE(x && scales && gx && gscales && rows >= 0 && C > 0 && C % 4 == 0, "gc_backward: needs C %% 4 == 0 and non-null x / scales / gx / gscales");
    LL_REQUIRE(g_out || g_q || g_rates, "gc_backward: no incoming gradient");
    LL_REQUIRE((means == nullptr) == (gmeans == nullptr), "gc_backward: means and gmeans go together");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(means) |
                 reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(g_out) | reinterpret_cast<uintptr_t>(g_q) |
                 reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(gscales) | reinterpret_cast<uintptr_t>(gmeans)) & 15) == 0,
               "gc_backward: tensors must be 16-byte aligned");
    if (rows == 0) return LL_OK;
    gc_backward_kernel<<<gc_grid(rows), kGcThreads, 0, as_stream(stream)>>>(x, scales, means, noise, rows, C, scale_bound, likelihood_bound, g_out, g_q, g_rates, gx, gscales, gmeans);
    LL_LAUNCH_CHECK();
    return LL_OK;
}
