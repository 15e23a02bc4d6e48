// Factorised-prior EntropyBottleneck rate: fused forward and backward (sm_100a).
//
// One launch replaces the ~60 tiny ATen / bmm launches of compressai==1.1.5
// EntropyBottleneck.forward (+ _likelihood, _logits_cumulative, LowerBound) wrapped by
// process_z_in / process_z_out and the -log q row sum of HRateFactorizedPrior.forward_help
// (lossyless/rates.py:534-554; hub/compressor.py:95,100-101).
//
// Layout: activations [rows, C] row-major; thread t owns channel c (adjacent threads ->
// adjacent channels -> coalesced 128 B warp accesses), keeps that channel's 58 derived
// parameters in registers and walks a slab of rows.  Per element: 2 x (66 FMA + 12 tanh) +
// 2 sigmoid + log; the kernel is SFU/FMA-bound, not HBM-bound.
#include "ll_common.cuh"

namespace ll {

// packed parameter rows (float32 [58, C]); M* = softplus(matrix), F* = tanh(factor)
enum : int {
    kM0 = 0, kB0 = 3, kF0 = 6,
    kM1 = 9, kB1 = 18, kF1 = 21,
    kM2 = 24, kB2 = 33, kF2 = 36,
    kM3 = 39, kB3 = 48, kF3 = 51,
    kM4 = 54, kB4 = 57,
    kNP = LL_EB_NPARAM
};
__device__ constexpr int kMOff[4] = {kM0, kM1, kM2, kM3};
__device__ constexpr int kBOff[4] = {kB0, kB1, kB2, kB3};
__device__ constexpr int kFOff[4] = {kF0, kF1, kF2, kF3};

// tanh and sigmoid from ex2.approx and ONE approximate reciprocal per group of three (relative error of either MUFU
// operation 2^-22): the three gated units of a layer share a reciprocal -- 1/(d0 d1 d2), then back to 1/d_j with four
// multiplications -- so a layer costs 3 EX2 + 1 RCP = 4 MUFU operations instead of 6, and a likelihood 36 instead of 53
// (the kernel is bound by the quarter-rate special-function pipe: 24 tanh + 2 sigmoid + 1 log per element, DESIGN.md
// 4.4).  Absolute error <= ~3e-7, the rounding level of the float32 activations they feed.  The full-precision
// libdevice versions cost ~25 instructions apiece and made the kernel 3x slower.
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// t_j = tanh(a_j) = 1 - 2 / (1 + e^{2 a_j}).  The exponent is clamped at 2^40 (tanh is 1 to 1e-12 there) so that the
// product of the three denominators stays finite.
__device__ __forceinline__ void tanh3(const float (&a)[3], float (&t)[3])
{
    float d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) d[j] = 1.f + ex2_approx(fminf(a[j] * 2.8853900817779268f, 40.f));
    const float p01 = d[0] * d[1];
    const float r = rcp_approx(p01 * d[2]);
    const float r01 = r * d[2];
    t[0] = fmaf(-2.f, r01 * d[1], 1.f);
    t[1] = fmaf(-2.f, r01 * d[0], 1.f);
    t[2] = fmaf(-2.f, r * p01, 1.f);
}

// logits_cumulative for one scalar; keeps gated activations g[k][j] and tanh values t[k][j]
struct Trace {
    float g[4][3];
    float t[4][3];
};

__device__ __forceinline__ float logits_fwd(const float (&P)[kNP], float v, Trace &tr)
{
    float a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = fmaf(P[kM0 + j], v, P[kB0 + j]);
    tanh3(a, tr.t[0]);
#pragma unroll
    for (int j = 0; j < 3; ++j) tr.g[0][j] = fmaf(P[kF0 + j], tr.t[0][j], a[j]);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[j] = P[kBOff[k] + j];
#pragma unroll
            for (int i = 0; i < 3; ++i) a[j] = fmaf(P[kMOff[k] + 3 * j + i], tr.g[k - 1][i], a[j]);
        }
        tanh3(a, tr.t[k]);
#pragma unroll
        for (int j = 0; j < 3; ++j) tr.g[k][j] = fmaf(P[kFOff[k] + j], tr.t[k][j], a[j]);
    }
    float out = P[kB4];
#pragma unroll
    for (int i = 0; i < 3; ++i) out = fmaf(P[kM4 + i], tr.g[3][i], out);
    return out;
}

// Forward-only logits_cumulative in the EXPONENT-SCALED domain: with c = 2 log2(e), the quantities carried from layer to
// layer are a^ = c a and g^ = c g, computed from pre-scaled parameters (B^ = c B, F^ = c F, M unchanged, last layer M^4 =
// M4 / c: packed rows [kNP, 2 kNP) of the parameter array).  tanh(a) = 1 - 2 / (1 + 2^{a^}) then needs no multiply in
// front of the ex2: 3 FMUL per layer and chain (24 per element) drop out of a kernel that is bound by instruction issue
// (ncu: 63 % issue-active at 360 warp instructions per element).  Same function up to one rounding of the parameters.
__device__ __forceinline__ void tanh3_scaled(const float (&as)[3], float (&t)[3])
{
    float d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) d[j] = 1.f + ex2_approx(fminf(as[j], 40.f));
    const float p01 = d[0] * d[1];
    const float r = rcp_approx(p01 * d[2]);
    const float r01 = r * d[2];
    t[0] = fmaf(-2.f, r01 * d[1], 1.f);
    t[1] = fmaf(-2.f, r01 * d[0], 1.f);
    t[2] = fmaf(-2.f, r * p01, 1.f);
}

__device__ __forceinline__ float logits_fwd_scaled(const float (&P)[kNP], float vs)  // vs = c * v; P = the scaled set
{
    float a[3], g[3], t[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = fmaf(P[kM0 + j], vs, P[kB0 + j]);
    tanh3_scaled(a, t);
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = fmaf(P[kF0 + j], t[j], a[j]);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[j] = P[kBOff[k] + j];
#pragma unroll
            for (int i = 0; i < 3; ++i) a[j] = fmaf(P[kMOff[k] + 3 * j + i], g[i], a[j]);
        }
        tanh3_scaled(a, t);
#pragma unroll
        for (int j = 0; j < 3; ++j) g[j] = fmaf(P[kFOff[k] + j], t[j], a[j]);
    }
    float out = P[kB4];
#pragma unroll
    for (int i = 0; i < 3; ++i) out = fmaf(P[kM4 + i], g[i], out);
    return out;
}

// sigmoid(x0), sigmoid(x1) with one reciprocal: s_j = 1 / (1 + e^{-x_j}); exponents clamped at 2^60 (sigmoid < 1e-18,
// far below the likelihood bound) so that the product of the two denominators stays finite
__device__ __forceinline__ void sigmoid2(float x0, float x1, float &s0, float &s1)
{
    const float d0 = 1.f + ex2_approx(fminf(-1.4426950408889634f * x0, 60.f));
    const float d1 = 1.f + ex2_approx(fminf(-1.4426950408889634f * x1, 60.f));
    const float r = rcp_approx(d0 * d1);
    s0 = r * d1;
    s1 = r * d0;
}

// accumulates d(out)/d(params) * gl into G and returns d(out)/dv * gl
__device__ __forceinline__ float logits_bwd(const float (&P)[kNP], float v, const Trace &tr, float gl, float (&G)[kNP])
{
    float gg[3];
    G[kB4] += gl;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        G[kM4 + i] = fmaf(gl, tr.g[3][i], G[kM4 + i]);
        gg[i] = gl * P[kM4 + i];
    }
#pragma unroll
    for (int k = 3; k >= 1; --k) {
        float ga[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float t = tr.t[k][j];
            G[kFOff[k] + j] = fmaf(gg[j], t, G[kFOff[k] + j]);
            ga[j] = gg[j] * fmaf(P[kFOff[k] + j], 1.f - t * t, 1.f);
            G[kBOff[k] + j] += ga[j];
        }
        float gprev[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                G[kMOff[k] + 3 * j + i] = fmaf(ga[j], tr.g[k - 1][i], G[kMOff[k] + 3 * j + i]);
                gprev[i] = fmaf(ga[j], P[kMOff[k] + 3 * j + i], gprev[i]);
            }
#pragma unroll
        for (int i = 0; i < 3; ++i) gg[i] = gprev[i];
    }
    float gv = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float t = tr.t[0][j];
        G[kF0 + j] = fmaf(gg[j], t, G[kF0 + j]);
        const float ga = gg[j] * fmaf(P[kF0 + j], 1.f - t * t, 1.f);
        G[kB0 + j] += ga;
        G[kM0 + j] = fmaf(ga, v, G[kM0 + j]);
        gv = fmaf(ga, P[kM0 + j], gv);
    }
    return gv;
}

__global__ void eb_pack_params_kernel(const float *m0, const float *m1, const float *m2, const float *m3, const float *m4,
                                      const float *b0, const float *b1, const float *b2, const float *b3, const float *b4,
                                      const float *f0, const float *f1, const float *f2, const float *f3, int C,
                                      float *__restrict__ packed)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    auto softplus = [](float x) { return x > 20.f ? x : log1pf(expf(x)); };  // torch.nn.functional.softplus
    const float *M[4] = {m0, m1, m2, m3};
    const float *B[4] = {b0, b1, b2, b3};
    const float *F[4] = {f0, f1, f2, f3};
    for (int j = 0; j < 3; ++j) packed[(kM0 + j) * C + c] = softplus(m0[c * 3 + j]);
    for (int k = 1; k < 4; ++k)
        for (int e = 0; e < 9; ++e) packed[(kMOff[k] + e) * C + c] = softplus(M[k][c * 9 + e]);
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 3; ++j) {
            packed[(kBOff[k] + j) * C + c] = B[k][c * 3 + j];
            packed[(kFOff[k] + j) * C + c] = tanhf(F[k][c * 3 + j]);
        }
    for (int i = 0; i < 3; ++i) packed[(kM4 + i) * C + c] = softplus(m4[c * 3 + i]);
    packed[kB4 * C + c] = b4[c];
    // rows [kNP, 2 kNP): the exponent-scaled set of logits_fwd_scaled
    const float cs = 2.8853900817779268f;
    float *ps = packed + (size_t)kNP * C;
    for (int j = 0; j < 3; ++j) ps[(kM0 + j) * C + c] = packed[(kM0 + j) * C + c];
    for (int k = 1; k < 4; ++k)
        for (int e = 0; e < 9; ++e) ps[(kMOff[k] + e) * C + c] = packed[(kMOff[k] + e) * C + c];
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 3; ++j) {
            ps[(kBOff[k] + j) * C + c] = cs * packed[(kBOff[k] + j) * C + c];
            ps[(kFOff[k] + j) * C + c] = cs * packed[(kFOff[k] + j) * C + c];
        }
    for (int i = 0; i < 3; ++i) ps[(kM4 + i) * C + c] = packed[(kM4 + i) * C + c] / cs;
    ps[kB4 * C + c] = packed[kB4 * C + c];
}

// Chain rule from the derived parameters back to the raw ones, all 14 tensors in ONE launch (it was ~60 tiny torch ops
// per backward call, 2.6 ms of host time per training step): g_matrix = g_softplus * sigmoid(matrix), g_factor = g_tanh *
// (1 - tanh(factor)^2), g_bias as is; outputs in the parameters' own layouts ([C, f_out, f_in] / [C, f_out, 1]).
__global__ void eb_unpack_grads_kernel(const float *__restrict__ gparam, const float *m0, const float *m1, const float *m2,
                                       const float *m3, const float *m4, const float *f0, const float *f1, const float *f2,
                                       const float *f3, int C, float *gm0, float *gm1, float *gm2, float *gm3, float *gm4, float *gb0,
                                       float *gb1, float *gb2, float *gb3, float *gb4, float *gf0, float *gf1, float *gf2, float *gf3)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    auto dsoftplus = [](float x) { return 1.f / (1.f + expf(-x)); };
    auto dtanh = [](float x) { const float t = tanhf(x); return 1.f - t * t; };
    const float *M[4] = {m0, m1, m2, m3};
    const float *F[4] = {f0, f1, f2, f3};
    float *GM[4] = {gm0, gm1, gm2, gm3};
    float *GB[4] = {gb0, gb1, gb2, gb3};
    float *GF[4] = {gf0, gf1, gf2, gf3};
    for (int j = 0; j < 3; ++j) gm0[c * 3 + j] = gparam[(kM0 + j) * C + c] * dsoftplus(m0[c * 3 + j]);
    for (int k = 1; k < 4; ++k)
        for (int e = 0; e < 9; ++e) GM[k][c * 9 + e] = gparam[(kMOff[k] + e) * C + c] * dsoftplus(M[k][c * 9 + e]);
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 3; ++j) {
            GB[k][c * 3 + j] = gparam[(kBOff[k] + j) * C + c];
            GF[k][c * 3 + j] = gparam[(kFOff[k] + j) * C + c] * dtanh(F[k][c * 3 + j]);
        }
    for (int i = 0; i < 3; ++i) gm4[c * 3 + i] = gparam[(kM4 + i) * C + c] * dsoftplus(m4[c * 3 + i]);
    gb4[c] = gparam[kB4 * C + c];
}

// compressai EntropyBottleneck.loss() (the "coder" optimiser's objective, lossyless/rates.py:426 aux_loss):
//   sum_c sum_k | logits_cumulative(quantiles[c, k]) - target[k] |   with the parameters detached,
// forward and the gradient w.r.t. the quantiles in one launch (thread per channel, 3 points each): ~90 tiny torch
// launches of forward + autograd otherwise.
__global__ void eb_aux_loss_kernel(const float *__restrict__ quantiles, const float *__restrict__ packed, const float *__restrict__ target,
                                   int C, float *__restrict__ loss, float *__restrict__ g_quantiles)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    if (c < C) {
        float P[kNP], G[kNP];
#pragma unroll
        for (int i = 0; i < kNP; ++i) P[i] = packed[i * C + c], G[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = quantiles[c * 3 + k];
            Trace tr;
            const float d = logits_fwd(P, v, tr) - target[k];
            part += fabsf(d);
            const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            g_quantiles[c * 3 + k] = logits_bwd(P, v, tr, sg, G);
        }
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, part);
}

constexpr int kEbThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kEbThreads)
eb_forward_kernel(const T *__restrict__ z, int64_t rows, int C, int rows_per_cta, const float *__restrict__ noise,
                  const float *__restrict__ bias, const float *__restrict__ exp_scale,
                  const float *__restrict__ medians, const float *__restrict__ packed, float bound,
                  float *__restrict__ z_hat, float *__restrict__ q_out, float *__restrict__ rates)
{
    const int c = blockIdx.x * kEbThreads + threadIdx.x;
    const bool live = c < C;
    const int cc = live ? c : C - 1;
    float P[kNP];  // the exponent-scaled parameter set (second half of the packed array)
#pragma unroll
    for (int i = 0; i < kNP; ++i) P[i] = packed[(size_t)(kNP + i) * C + cc];
    const float b = bias ? bias[cc] : 0.f, es = exp_scale ? exp_scale[cc] : 1.f, med = medians[cc];
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    // two rows per iteration: four independent logits chains (lower / upper of each row) in flight per thread
    for (int64_t r = r0; r < r1; r += 2) {
        const bool two = r + 1 < r1;
        const int64_t idx0 = r * C + cc, idx1 = two ? idx0 + C : idx0;
        float x[2] = {to_float<T>(z[idx0]), to_float<T>(z[idx1])};
        float nz[2] = {0.f, 0.f};
        if (noise) nz[0] = noise[idx0], nz[1] = noise[idx1];
        float v[2], nl[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (bias) x[u] = __fmul_rn(__fadd_rn(x[u], b), es);  // process_z_in
            v[u] = noise ? x[u] + nz[u] : __fadd_rn(rintf(__fsub_rn(x[u], med)), med);  // quantize "noise" / "dequantize"
        }
        float lower[2], upper[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            lower[u] = logits_fwd_scaled(P, 2.8853900817779268f * (v[u] - 0.5f));
            upper[u] = logits_fwd_scaled(P, 2.8853900817779268f * (v[u] + 0.5f));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float sum = lower[u] + upper[u];
            const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
            float su, sl;
            sigmoid2(sgn * upper[u], sgn * lower[u], su, sl);
            const float lik = fabsf(su - sl);
            const float q = fmaxf(lik, bound);  // LowerBound
            nl[u] = live && (u == 0 || two) ? -__logf(q) : 0.f;
            if (live && (u == 0 || two)) {
                const int64_t idx = u ? idx1 : idx0;
                z_hat[idx] = bias ? __fsub_rn(__fdiv_rn(v[u], es), b) : v[u];  // process_z_out
                if (q_out) q_out[idx] = q;
            }
        }
        if (rates) {
            nl[0] = warp_sum(nl[0]);
            nl[1] = warp_sum(nl[1]);
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(&rates[r], nl[0]);
                if (two) atomicAdd(&rates[r + 1], nl[1]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kEbThreads)
eb_backward_kernel(const T *__restrict__ z, int64_t rows, int C, int rows_per_cta, const float *__restrict__ noise,
                   const float *__restrict__ bias, const float *__restrict__ exp_scale,
                   const float *__restrict__ medians, const float *__restrict__ packed, float bound,
                   const float *__restrict__ g_rates, const float *__restrict__ g_q,
                   const float *__restrict__ g_zhat, float *__restrict__ gz, float *__restrict__ gparam,
                   float *__restrict__ g_bias, float *__restrict__ g_es)
{
    const int c = blockIdx.x * kEbThreads + threadIdx.x;
    if (c >= C) return;
    float P[kNP], G[kNP];
#pragma unroll
    for (int i = 0; i < kNP; ++i) {
        P[i] = packed[i * C + c];
        G[i] = 0.f;
    }
    const float b = bias ? bias[c] : 0.f, es = exp_scale ? exp_scale[c] : 1.f;
    const float med = noise ? 0.f : medians[c];
    float acc_b = 0.f, acc_es = 0.f;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r1 = min(rows, r0 + rows_per_cta);
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t idx = r * C + c;
        const float zf = to_float<T>(z[idx]);
        const float zb = __fadd_rn(zf, b);
        const float x = bias ? __fmul_rn(zb, es) : zf;  // the forward's own rounding (matters for round() in eval mode)
        // training: v = x + noise (dv/dx = 1); eval: v = round(x - med) + med (dv/dx = 0, as torch.round)
        const float v = noise ? x + noise[idx] : __fadd_rn(rintf(__fsub_rn(x, med)), med);
        Trace trl, tru;
        const float lower = logits_fwd(P, v - 0.5f, trl);
        const float upper = logits_fwd(P, v + 0.5f, tru);
        const float sum = lower + upper;
        const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
        float su, sl;
        sigmoid2(sgn * upper, sgn * lower, su, sl);
        const float diff = su - sl;
        const float lik = fabsf(diff);
        const float q = fmaxf(lik, bound);
        // rate = -log q ;  LowerBound passes the gradient where lik >= bound or grad < 0
        // (+ a gradient arriving directly on the likelihood output, e.g. -log q(s) taken outside the kernel)
        float g_lik = (g_rates ? -g_rates[r] / q : 0.f) + (g_q ? g_q[idx] : 0.f);
        if (!(lik >= bound || g_lik < 0.f)) g_lik = 0.f;
        const float sd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float gu = g_lik * sd * sgn * su * (1.f - su);
        const float gl = -g_lik * sd * sgn * sl * (1.f - sl);
        float gv = logits_bwd(P, v + 0.5f, tru, gu, G) + logits_bwd(P, v - 0.5f, trl, gl, G);
        if (!noise) gv = 0.f;
        if (g_zhat) {
            const float gh = g_zhat[idx];
            if (bias) {
                if (noise) gv += gh / es;     // z_hat = v / es - b
                acc_es -= gh * v / (es * es);
                acc_b -= gh;
            } else if (noise) {
                gv += gh;
            }
        }
        if (bias) {
            gz[idx] = gv * es;  // x = (z + b) * es
            acc_b += gv * es;
            acc_es += gv * zb;
        } else {
            gz[idx] = gv;
        }
    }
#pragma unroll
    for (int i = 0; i < kNP; ++i) atomicAdd(&gparam[i * C + c], G[i]);
    if (bias) {
        atomicAdd(&g_bias[c], acc_b);
        atomicAdd(&g_es[c], acc_es);
    }
}

// One wave: `per_sm` = the kernel's resident CTAs per SM (occupancy query), so that the grid is 148 x per_sm CTAs and
// no SM waits for a second, partially filled wave; at least 8 rows per slab to amortise the parameter load.
static int eb_grid(int64_t rows, int C, int per_sm, dim3 &grid, int &rows_per_cta)
{
    const int cx = (C + kEbThreads - 1) / kEbThreads;
    int64_t slabs = ((int64_t)per_sm * num_sms()) / cx;
    if (slabs < 1) slabs = 1;
    int64_t rpc = (rows + slabs - 1) / slabs;
    rpc = (rpc + 1) & ~(int64_t)1;  // the forward kernel walks row pairs
    if (rpc < 8) rpc = 8;
    slabs = (rows + rpc - 1) / rpc;
    if (slabs > 65535) {
        slabs = 65535;
        rpc = ((rows + slabs - 1) / slabs + 1) & ~(int64_t)1;
        slabs = (rows + rpc - 1) / rpc;
    }
    rows_per_cta = (int)rpc;
    grid = dim3((unsigned)cx, (unsigned)slabs, 1);
    return LL_OK;
}

template <typename K>
static int resident_ctas(K kernel)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kEbThreads, 0) != cudaSuccess || n < 1) n = 4;
    return n;
}

}  // namespace ll

using namespace ll;

extern "C" int ll_eb_pack_params(const float *const *m, const float *const *b, const float *const *f, int C,
                                 float *packed, ll_stream_t stream)
{
    LL_REQUIRE(m && b && f && packed && C > 0, "eb_pack_params: bad arguments");
    for (int i = 0; i < 5; ++i) LL_REQUIRE(m[i] && b[i], "eb_pack_params: null matrix/bias %d", i);
    for (int i = 0; i < 4; ++i) LL_REQUIRE(f[i], "eb_pack_params: null factor %d", i);
    eb_pack_params_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(m[0], m[1], m[2], m[3], m[4], b[0], b[1], b[2], b[3], b[4], f[0], f[1], f[2], f[3], C, packed);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_eb_forward(const void *z, int z_dtype, int64_t rows, int C, const float *noise, const float *bias,
                             const float *exp_scale, const float *medians, const float *packed, float bound,
                             float *z_hat, float *q, float *rates, ll_stream_t stream)
{
    LL_REQUIRE(z && medians && packed && z_hat, "eb_forward: null pointer");
    LL_REQUIRE((bias == nullptr) == (exp_scale == nullptr), "eb_forward: bias and exp_scale go together");
    LL_REQUIRE(rows >= 0 && C > 0, "eb_forward: bad sizes");
    if (rows == 0) return LL_OK;
    dim3 grid;
    int rpc;
    static const int per_sm = resident_ctas(eb_forward_kernel<float>);
    eb_grid(rows, C, per_sm, grid, rpc);
    cudaStream_t s = as_stream(stream);
    switch (z_dtype) {
        case LL_F32: eb_forward_kernel<float><<<grid, kEbThreads, 0, s>>>((const float *)z, rows, C, rpc, noise, bias, exp_scale, medians, packed, bound, z_hat, q, rates); break;
        case LL_F16: eb_forward_kernel<__half><<<grid, kEbThreads, 0, s>>>((const __half *)z, rows, C, rpc, noise, bias, exp_scale, medians, packed, bound, z_hat, q, rates); break;
        case LL_BF16: eb_forward_kernel<__nv_bfloat16><<<grid, kEbThreads, 0, s>>>((const __nv_bfloat16 *)z, rows, C, rpc, noise, bias, exp_scale, medians, packed, bound, z_hat, q, rates); break;
        default: return fail(LL_ERR_INVALID, "eb_forward: unknown dtype %d", z_dtype);
    }
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_eb_backward(const void *z, int z_dtype, int64_t rows, int C, const float *noise, const float *bias,
                              const float *exp_scale, const float *medians, const float *packed, float bound,
                              const float *g_rates, const float *g_q, const float *g_zhat, float *gz, float *gparam,
                              float *g_bias, float *g_exp_scale, ll_stream_t stream)
{
    LL_REQUIRE(z && packed && gz && gparam, "eb_backward: null pointer");
    LL_REQUIRE(noise || medians, "eb_backward: eval mode (noise == NULL) needs the medians");
    LL_REQUIRE(g_rates || g_q || g_zhat, "eb_backward: no incoming gradient");
    LL_REQUIRE((bias == nullptr) == (exp_scale == nullptr), "eb_backward: bias and exp_scale go together");
    LL_REQUIRE(!bias || (g_bias && g_exp_scale), "eb_backward: g_bias / g_exp_scale required with bias");
    LL_REQUIRE(rows >= 0 && C > 0, "eb_backward: bad sizes");
    if (rows == 0) return LL_OK;
    dim3 grid;
    int rpc;
    static const int per_sm = resident_ctas(eb_backward_kernel<float>);
    eb_grid(rows, C, per_sm, grid, rpc);
    cudaStream_t s = as_stream(stream);
    switch (z_dtype) {
        case LL_F32: eb_backward_kernel<float><<<grid, kEbThreads, 0, s>>>((const float *)z, rows, C, rpc, noise, bias, exp_scale, medians, packed, bound, g_rates, g_q, g_zhat, gz, gparam, g_bias, g_exp_scale); break;
        case LL_F16: eb_backward_kernel<__half><<<grid, kEbThreads, 0, s>>>((const __half *)z, rows, C, rpc, noise, bias, exp_scale, medians, packed, bound, g_rates, g_q, g_zhat, gz, gparam, g_bias, g_exp_scale); break;
        case LL_BF16: eb_backward_kernel<__nv_bfloat16><<<grid, kEbThreads, 0, s>>>((const __nv_bfloat16 *)z, rows, C, rpc, noise, bias, exp_scale, medians, packed, bound, g_rates, g_q, g_zhat, gz, gparam, g_bias, g_exp_scale); break;
        default: return fail(LL_ERR_INVALID, "eb_backward: unknown dtype %d", z_dtype);
    }
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_eb_unpack_grads(const float *gparam, const float *const *m, const float *const *f, int C, float *const *g_m,
                                  float *const *g_b, float *const *g_f, ll_stream_t stream)
{
    LL_REQUIRE(gparam && m && f && g_m && g_b && g_f && C > 0, "eb_unpack_grads: bad arguments");
    for (int i = 0; i < 5; ++i) LL_REQUIRE(m[i] && g_m[i] && g_b[i], "eb_unpack_grads: null matrix / gradient %d", i);
    for (int i = 0; i < 4; ++i) LL_REQUIRE(f[i] && g_f[i], "eb_unpack_grads: null factor / gradient %d", i);
    eb_unpack_grads_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(gparam, m[0], m[1], m[2], m[3], m[4], f[0], f[1], f[2], f[3], C,
                                                                        g_m[0], g_m[1], g_m[2], g_m[3], g_m[4], g_b[0], g_b[1], g_b[2],
                                                                        g_b[3], g_b[4], g_f[0], g_f[1], g_f[2], g_f[3]);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_eb_aux_loss(const float *quantiles, const float *packed, const float *target, int C, float *loss,
                              float *g_quantiles, ll_stream_t stream)
{
    LL_REQUIRE(quantiles && packed && target && loss && g_quantiles && C > 0, "eb_aux_loss: bad arguments");
    eb_aux_loss_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(quantiles, packed, target, C, loss, g_quantiles);
    LL_LAUNCH_CHECK();
    return LL_OK;
}
