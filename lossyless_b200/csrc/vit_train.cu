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
#include <algorithm>

#include "ll_common.cuh"

namespace ll {

// ------------------------------------------------------------------------- transpose (2-byte elements)
// 64 x 64 tiles, 4-byte global accesses on both sides (a warp moves one 128-byte row segment per access).
// Needs C even and Rpad even (true for every caller: widths and padded token counts are multiples of 64);
// odd shapes take the scalar kernel below.
template <typename T, bool kColsum>
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int64_t R, int C, int64_t Rpad,
                     float *__restrict__ colsum)
{
    pdl_wait();  // programmatic dependent launch: nothing of global memory is touched before the predecessor is complete
    pdl_launch_dependents();
    __shared__ uint16_t tile[64][66];
    const int64_t r0 = (int64_t)blockIdx.y * 64;
    const int c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int64_t r = r0 + ty + 8 * k;
        const int c = c0 + 2 * tx;
        uint32_t v = 0;
        if (r < R && c < C) v = *reinterpret_cast<const uint32_t *>(src + r * C + c);  // C even: c + 1 < C too
        *reinterpret_cast<uint32_t *>(&tile[ty + 8 * k][2 * tx]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int cl = ty + 8 * k;
        const int c = c0 + cl;
        const int64_t r = r0 + 2 * tx;
        if (c < C && r < Rpad) {
            const uint32_t v = (uint32_t)tile[2 * tx][cl] | ((uint32_t)tile[2 * tx + 1][cl] << 16);
            *reinterpret_cast<uint32_t *>(dst + (int64_t)c * Rpad + r) = v;  // Rpad even: r + 1 < Rpad too
        }
        if (kColsum) {
            // bias gradient on the way: the warp holds column c's 64 rows of this tile (zero beyond R)
            const T *col0 = reinterpret_cast<const T *>(&tile[2 * tx][cl]), *col1 = reinterpret_cast<const T *>(&tile[2 * tx + 1][cl]);
            const float part = warp_sum(to_float<T>(*col0) + to_float<T>(*col1));
            if (tx == 0 && c < C) atomicAdd(colsum + c, part);
        }
    }
}

__global__ void __launch_bounds__(256)
transpose_pad_scalar_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int64_t R, int C, int64_t Rpad)
{
    __shared__ uint16_t tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
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
// out[c] += sum_r src[r, c].  A CTA covers 256 columns x rows_per_cta rows (a multiple of 64 sized for about four
// CTAs per SM instead of thousands of 32-row CTAs whose life was mostly their reduction and their 256 atomics: +17 % on
// the 768-column shape, the wide ones were at the HBM roofline already):
// thread = (32 column groups of 8) x (8 row
// lanes), one 16-byte load per thread and row, eight rows in flight; the 8 row lanes are combined through shared
// memory and each CTA adds its 256 partial sums with one atomic per column.  (The first version read 2 bytes
// per thread: 1.2 TB/s.)  Needs C % 8 == 0 and a 16-byte aligned base; other shapes take the scalar kernel.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T *__restrict__ src, int64_t R, int C, int rows_per_cta, float *__restrict__ out);

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

template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T *__restrict__ src, int64_t R, int C, int rows_per_cta, float *__restrict__ out)
{
    pdl_wait();  // programmatic dependent launch: nothing of global memory is touched before the predecessor is complete
    pdl_launch_dependents();
    __shared__ float part[8][256 + 8];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + cg * 8;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta;
    const int64_t r_end = min(R, r_begin + rows_per_cta);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const uint32_t never = rows_per_cta < 0 ? 0xFFFFFFFFu : 0u;  // always 0; unknown to the compiler
    if (c0 < C) {
        int64_t r = r_begin + rl;
        const T *p = src + r * C + c0;
        const int64_t step = (int64_t)8 * C;  // this thread's next row
        for (; r + 56 < r_end; r += 64, p += 8 * step) {  // eight independent 16-byte loads in flight per thread
            uint4 raw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)  // volatile: issued back to back (the plain loads were scheduled two at a time)
                asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(raw[j].x), "=r"(raw[j].y), "=r"(raw[j].z), "=r"(raw[j].w) : "l"(p + j * step));
            // ptxas would sink the loads between the sums (two or three in flight, the rest stalled behind the first
            // sum's wait on an in-order pipe); a value it cannot prove zero ties the first sum to the LAST load
            raw[0].x ^= raw[7].w & never;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const T *e = reinterpret_cast<const T *>(&raw[j]);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += to_float<T>(e[k]);
            }
        }
        for (; r < r_end; r += 8, p += step) {
            float a[8];
            load8<T>(p, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += a[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[rl][cg * 8 + k] = acc[k];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < C) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += part[j][threadIdx.x];
        atomicAdd(out + c, sum);
    }
}

// ------------------------------------------------------------------------- token embedding (training forward / backward)
// VisionTransformer.forward: x = cat([class_embedding, conv1(x)], dim = 1) + positional_embedding.
//   tok[b, 0, :] = cls + pos[0],  tok[b, s, :] = patch[b, s - 1, :] + pos[s]      (one 16-bit rounding per element, as the
// 16-bit torch add; replaces a CatArrayBatchedCopy and an elementwise add: 167 us -> one pass over the tensor)
template <typename T>
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const T *__restrict__ patch, const T *__restrict__ cls, const T *__restrict__ pos, T *__restrict__ tok,
                    int64_t B, int S, int W)
{
    pdl_wait();
    pdl_launch_dependents();
    const int w8 = W / 8;
    const int64_t n = B * S * w8, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = (int)(i % w8) * 8;
        const int64_t row = i / w8;
        const int s = (int)(row % S);
        const int64_t b = row / S;
        float a[8], p[8];
        load8<T>(s == 0 ? cls + c : patch + (b * (S - 1) + (s - 1)) * W + c, a);
        load8<T>(pos + (int64_t)s * W + c, p);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] + p[k];
        store8<T>(tok + row * W + c, a);
    }
}

// Its adjoint: g_patch[b, s - 1, :] = g_tok[b, s, :] (the contiguous operand of conv1's weight-gradient GEMM),
// g_pos[s, :] += sum_b g_tok[b, s, :], g_cls += sum_b g_tok[b, 0, :] (float32) -- g_tok is read once.  A CTA covers 256
// columns of one token position over a slice of the batch: thread = (32 column groups of 8) x (8 batch lanes).
template <typename T>
__global__ void __launch_bounds__(256)
embed_backward_kernel(const T *__restrict__ g_tok, T *__restrict__ g_patch, float *__restrict__ g_pos, float *__restrict__ g_cls,
                      int64_t B, int S, int W, int batch_per_cta)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float part[8][256 + 8];
    const int cg = threadIdx.x & 31, bl = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + cg * 8, s = blockIdx.y;
    const int64_t b_begin = (int64_t)blockIdx.z * batch_per_cta, b_end = min(B, b_begin + batch_per_cta);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < W) {
        for (int64_t b = b_begin + bl; b < b_end; b += 8) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(g_tok + (b * S + s) * W + c0);
            if (s > 0) *reinterpret_cast<uint4 *>(g_patch + (b * (S - 1) + (s - 1)) * W + c0) = raw;
            const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += to_float<T>(e[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[bl][cg * 8 + k] = acc[k];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < W) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += part[j][threadIdx.x];
        atomicAdd(g_pos + (int64_t)s * W + c, sum);
        if (s == 0) atomicAdd(g_cls + c, sum);
    }
}

// ------------------------------------------------------------------------- QuickGELU
// 8 elements (16 bytes) per thread per step; the tail (n % 8) is handled by the last thread scalar-wise.
template <typename T>
__global__ void __launch_bounds__(256)
quickgelu_forward_kernel(const T *__restrict__ u, T *__restrict__ f, int64_t n)
{
    pdl_wait();  // programmatic dependent launch: nothing of global memory is touched before the predecessor is complete
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = n / 8;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n8; i += 2 * stride) {  // two independent 16-byte loads in flight per thread
        float x[8], y[8];
        load8<T>(u + 8 * i, x);
        load8<T>(u + 8 * (i + stride), y);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = x[k] / (1.f + __expf(-1.702f * x[k])), y[k] = y[k] / (1.f + __expf(-1.702f * y[k]));
        store8<T>(f + 8 * i, x);
        store8<T>(f + 8 * (i + stride), y);
    }
    for (; i < n8; i += stride) {
        float x[8];
        load8<T>(u + 8 * i, x);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = x[k] / (1.f + __expf(-1.702f * x[k]));
        store8<T>(f + 8 * i, x);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = 8 * n8; i < n; ++i) {
            const float x = to_float<T>(u[i]);
            f[i] = from_float<T>(x / (1.f + __expf(-1.702f * x)));
        }
}

template <typename T>
__global__ void __launch_bounds__(256)
quickgelu_backward_kernel(const T *__restrict__ u, const T *__restrict__ g_f, T *__restrict__ g_u, int64_t n)
{
    pdl_wait();  // programmatic dependent launch: nothing of global memory is touched before the predecessor is complete
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = n / 8;
    // d/dx [x s(1.702 x)] = s + 1.702 x s (1 - s)
    auto dgelu = [](float x, float g) {
        const float sg = 1.f / (1.f + __expf(-1.702f * x));
        return g * (sg + 1.702f * x * sg * (1.f - sg));
    };
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n8; i += 2 * stride) {  // four independent 16-byte loads in flight per thread
        float x[8], g[8], y[8], h[8];
        load8<T>(u + 8 * i, x);
        load8<T>(g_f + 8 * i, g);
        load8<T>(u + 8 * (i + stride), y);
        load8<T>(g_f + 8 * (i + stride), h);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = dgelu(x[k], g[k]), h[k] = dgelu(y[k], h[k]);
        store8<T>(g_u + 8 * i, g);
        store8<T>(g_u + 8 * (i + stride), h);
    }
    for (; i < n8; i += stride) {
        float x[8], g[8];
        load8<T>(u + 8 * i, x);
        load8<T>(g_f + 8 * i, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = dgelu(x[k], g[k]);
        store8<T>(g_u + 8 * i, g);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = 8 * n8; i < n; ++i) {
            const float x = to_float<T>(u[i]);
            const float sg = 1.f / (1.f + __expf(-1.702f * x));
            g_u[i] = from_float<T>(to_float<T>(g_f[i]) * (sg + 1.702f * x * sg * (1.f - sg)));
        }
}

// ------------------------------------------------------------------------- LayerNorm backward
// h = (x - mean) * rstd * gamma + beta (clip/model.py LayerNorm: fp32 statistics, eps 1e-5).
//   g_x = g_res + rstd * (g_h*gamma - mean_c(g_h*gamma) - xhat * mean_c(g_h*gamma*xhat))
//   g_gamma += sum_rows g_h * xhat,  g_beta += sum_rows g_h
// One warp per row (grid-stride); a lane owns 8 contiguous columns in each 256-column block.  The row's three inputs
// (x, g_h, g_res) stay PACKED in registers (9 x 16 bytes per lane, all loads in flight together) and are unpacked again
// in each of the four passes over the row -- the float copies of the first version took the kernel to 190 registers,
// one 256-thread CTA per SM, and 2.3 TB/s; at <= 128 registers two CTAs are resident.
template <typename T>
__device__ __forceinline__ void unpack8(const uint4 &raw, float *out)
{
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = to_float<T>(e[k]);
}

// Epilogue of layernorm_backward_kernel: the CTA's eight per-warp column sums meet in shared memory through plain stores
// and ONE coalesced atomic per column and CTA goes to global memory (thread t of the CTA owns columns b * 256 + t: a
// warp's 32 atomics fall into one 128-byte line).  Column c is parked at slot c + c / 32: the writers (lane l holds
// columns 8 l .. 8 l + 7) and the readers (consecutive columns) are both bank-conflict-free.
// (The first version used shared-memory float atomics -- compare-and-swap loops, 48 per thread: 35.5 -> 31.1 us per
// 25 600-row launch without them.  An intermediate layout that let every warp add to columns 32 bytes apart in GLOBAL
// memory made each 32-byte sector of the 6 KB target receive 2368 dependent read-modify-writes instead of 296: a quarter
// of ncu's warp samples sat behind EXIT waiting for them to drain, and the kernel took longer than its pass over the rows.)
constexpr int kLnSlots = 256 + 8;  // padded slots per 256-column block
template <int NB>
__device__ __forceinline__ void ln_bwd_combine(float (*s_part)[NB * kLnSlots], const float (&acc)[NB][8], float *__restrict__ dst, int lane, int warp)
{
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) s_part[warp][b * kLnSlots + lane * 8 + k + lane / 4] = acc[b][k];
    __syncthreads();
    const int t = warp * 32 + lane;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += s_part[w][b * kLnSlots + t + t / 32];
        atomicAdd(dst + b * 256 + t, sum);
    }
}

template <typename T, int NB>
__global__ void __launch_bounds__(256, NB <= 3 ? 2 : 1)
layernorm_backward_kernel(const T *__restrict__ x, const T *__restrict__ g_h, const float *__restrict__ gamma,
                          const T *__restrict__ g_res, T *__restrict__ g_x, float *__restrict__ g_gamma,
                          float *__restrict__ g_beta, int64_t rows)
{
    constexpr int W = NB * 256;
    // gamma permuted so that a lane's two float4 pieces of a block are conflict-free: [block][half][lane][4]
    __shared__ float4 s_gam[NB * 2 * 32];
    __shared__ float s_part[8][NB * kLnSlots];  // the warps' column sums, combined once per CTA (epilogue)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_wait();
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < NB * 64; i += 256) {
        const int b = i / 64, half = (i / 32) & 1, l = i & 31;
        s_gam[i] = *reinterpret_cast<const float4 *>(gamma + b * 256 + l * 8 + half * 4);
    }
    __syncthreads();
    float acc_g[NB][8], acc_b[NB][8];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc_g[b][k] = 0.f, acc_b[b][k] = 0.f;
    const int64_t warps_total = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += warps_total) {
        uint4 xr[NB], gr[NB], rr[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int64_t at = r * W + b * 256 + lane * 8;
            xr[b] = *reinterpret_cast<const uint4 *>(x + at);
            gr[b] = *reinterpret_cast<const uint4 *>(g_h + at);
            rr[b] = g_res != nullptr ? *reinterpret_cast<const uint4 *>(g_res + at) : make_uint4(0, 0, 0, 0);
        }
        float sum = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float xv[8];
            unpack8<T>(xr[b], xv);
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += xv[k];
        }
        const float mean = warp_sum(sum) * (1.f / W);
        float var = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float xv[8];
            unpack8<T>(xr[b], xv);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float d = xv[k] - mean;
                var += d * d;
            }
        }
        const float rstd = rsqrtf(warp_sum(var) * (1.f / W) + 1e-5f);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float xv[8], gv[8];
            unpack8<T>(xr[b], xv);
            unpack8<T>(gr[b], gv);
            const float4 ga = s_gam[(b * 2) * 32 + lane], gb = s_gam[(b * 2 + 1) * 32 + lane];
            const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = (xv[k] - mean) * rstd;
                acc_g[b][k] += gv[k] * xh;
                acc_b[b][k] += gv[k];
                const float d = gv[k] * gam[k];  // d xhat
                m1 += d;
                m2 += d * xh;
            }
        }
        m1 = warp_sum(m1) * (1.f / W);
        m2 = warp_sum(m2) * (1.f / W);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float xv[8], gv[8], res[8], g[8];
            unpack8<T>(xr[b], xv);
            unpack8<T>(gr[b], gv);
            unpack8<T>(rr[b], res);
            const float4 ga = s_gam[(b * 2) * 32 + lane], gb = s_gam[(b * 2 + 1) * 32 + lane];
            const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = (xv[k] - mean) * rstd;
                g[k] = rstd * (gv[k] * gam[k] - m1 - xh * m2) + res[k];
            }
            store8<T>(g_x + r * W + b * 256 + lane * 8, g);
        }
    }
    ln_bwd_combine<NB>(s_part, acc_g, g_gamma, lane, warp);
    __syncthreads();
    ln_bwd_combine<NB>(s_part, acc_b, g_beta, lane, warp);
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

// Tensor-core variant (default): the five products of the backward pass on mma.sync.m16n8k16 with fp32
// accumulation, S padded to 64.  Four warps; warp w owns rows 16w..16w+15 of every product's M dimension.
// Operands live in shared memory in both orientations ([s][d] and [d][s]; P and dS as [i][j] and [j][i],
// rounded to the 16-bit type as flash-attention backward does) so that every fragment is two plain 32-bit
// loads per register pair with the reduction index contiguous; the 72-element pitch makes them conflict-free.
template <typename T>
struct TrainMma;
template <>
struct TrainMma<__half> {
    static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
    {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <>
struct TrainMma<__nv_bfloat16> {
    static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
    {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

constexpr int kAmP = 72;  // elements per shared-memory row (144 B: rows g = 0..7 land on banks 4g + t)
constexpr size_t kAmSmem = 11 * 64 * kAmP * 2;

template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi)
{
    T v[2] = {from_float<T>(lo), from_float<T>(hi)};
    return *reinterpret_cast<uint32_t *>(v);
}

// acc[nt] (16 x 8 tiles, nt = 0..7) += A[m0.., :] * B^T with A = [m][k] and B = [n][k], k = 0..63
template <typename T>
__device__ __forceinline__ void warp_mma_64(float (&acc)[8][4], const T *A, const T *B, int m0, int g, int t)
{
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t *>(A + (m0 + g) * kAmP + 16 * kk + 2 * t);
        a[1] = *reinterpret_cast<const uint32_t *>(A + (m0 + g + 8) * kAmP + 16 * kk + 2 * t);
        a[2] = *reinterpret_cast<const uint32_t *>(A + (m0 + g) * kAmP + 16 * kk + 2 * t + 8);
        a[3] = *reinterpret_cast<const uint32_t *>(A + (m0 + g + 8) * kAmP + 16 * kk + 2 * t + 8);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(B + (8 * nt + g) * kAmP + 16 * kk + 2 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(B + (8 * nt + g) * kAmP + 16 * kk + 2 * t + 8);
            TrainMma<T>::mma(acc[nt], a, b0, b1);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
attention_backward_mma_kernel(const T *__restrict__ qkv, const T *__restrict__ g_out, T *__restrict__ g_qkv, int S, int H)
{
    extern __shared__ __align__(16) uint8_t ab_smem[];
    T *sQ = reinterpret_cast<T *>(ab_smem);
    T *sK = sQ + 64 * kAmP, *sV = sK + 64 * kAmP, *sG = sV + 64 * kAmP;          // [s][d]
    T *sQt = sG + 64 * kAmP, *sKt = sQt + 64 * kAmP, *sGt = sKt + 64 * kAmP;     // [d][s]
    T *sP = sGt + 64 * kAmP, *sPt = sP + 64 * kAmP, *sS = sPt + 64 * kAmP, *sSt = sS + 64 * kAmP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int img = blockIdx.x / H, head = blockIdx.x % H;
    const int64_t row0 = (int64_t)img * S;
    const int ld = 3 * H * kAbD, ldo = H * kAbD;
    const float scale = 0.125f;

    for (int i = tid; i < 64 * 32; i += 128) {
        const int r = i >> 5, w = i & 31;
        uint32_t q = 0, k = 0, v = 0, o = 0;
        if (r < S) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(qkv + (row0 + r) * ld + head * kAbD);
            q = src[w];
            k = src[w + (H * kAbD) / 2];
            v = src[w + H * kAbD];
            o = reinterpret_cast<const uint32_t *>(g_out + (row0 + r) * ldo + head * kAbD)[w];
        }
        *reinterpret_cast<uint32_t *>(sQ + r * kAmP + 2 * w) = q;
        *reinterpret_cast<uint32_t *>(sK + r * kAmP + 2 * w) = k;
        *reinterpret_cast<uint32_t *>(sV + r * kAmP + 2 * w) = v;
        *reinterpret_cast<uint32_t *>(sG + r * kAmP + 2 * w) = o;
        const T *qe = reinterpret_cast<const T *>(&q), *ke = reinterpret_cast<const T *>(&k), *oe = reinterpret_cast<const T *>(&o);
        sQt[(2 * w) * kAmP + r] = qe[0];
        sQt[(2 * w + 1) * kAmP + r] = qe[1];
        sKt[(2 * w) * kAmP + r] = ke[0];
        sKt[(2 * w + 1) * kAmP + r] = ke[1];
        sGt[(2 * w) * kAmP + r] = oe[0];
        sGt[(2 * w + 1) * kAmP + r] = oe[1];
    }
    __syncthreads();

    const int m0 = 16 * warp;
    {
        // scores and dP for rows m0 .. m0+15
        float sc[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f, dp[nt][e] = 0.f;
        warp_mma_64<T>(sc, sQ, sK, m0, g, t);
        warp_mma_64<T>(dp, sG, sV, m0, g, t);
        // element e of tile nt: row m0 + g + 8 * (e >> 1), column 8 nt + 2 t + (e & 1)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * nt + 2 * t + e;
                    float &x = sc[nt][2 * h + e];
                    x = col < S ? x * scale : -INFINITY;
                    mx = fmaxf(mx, x);
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float &x = sc[nt][2 * h + e];
                    x = __expf(x - mx);  // exp(-inf) = 0 for the padded columns
                    sum += x;
                }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = 1.f / sum;
            float delta = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    sc[nt][2 * h + e] *= inv;
                    delta += sc[nt][2 * h + e] * dp[nt][2 * h + e];
                }
            delta += __shfl_xor_sync(0xffffffffu, delta, 1);
            delta += __shfl_xor_sync(0xffffffffu, delta, 2);
            const int row = m0 + g + 8 * h;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int col = 8 * nt + 2 * t;
                const float p0 = sc[nt][2 * h], p1 = sc[nt][2 * h + 1];
                const float d0 = p0 * (dp[nt][2 * h] - delta) * scale, d1 = p1 * (dp[nt][2 * h + 1] - delta) * scale;
                *reinterpret_cast<uint32_t *>(sP + row * kAmP + col) = pack2<T>(p0, p1);
                *reinterpret_cast<uint32_t *>(sS + row * kAmP + col) = pack2<T>(d0, d1);
                sPt[col * kAmP + row] = from_float<T>(p0);
                sPt[(col + 1) * kAmP + row] = from_float<T>(p1);
                sSt[col * kAmP + row] = from_float<T>(d0);
                sSt[(col + 1) * kAmP + row] = from_float<T>(d1);
            }
        }
    }
    __syncthreads();

    // dQ = dS K, dK = dS^T Q, dV = P^T dO for rows m0 .. m0+15
    float acc[8][4];
    auto store = [&](int which) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = m0 + g + 8 * h;
            if (row < S) {
                T *dst = g_qkv + (row0 + row) * ld + which * H * kAbD + head * kAbD;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t *>(dst + 8 * nt + 2 * t) = pack2<T>(acc[nt][2 * h], acc[nt][2 * h + 1]);
            }
        }
    };
    auto clear = [&]() {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    };
    clear();
    warp_mma_64<T>(acc, sS, sKt, m0, g, t);
    store(0);
    clear();
    warp_mma_64<T>(acc, sSt, sQt, m0, g, t);
    store(1);
    clear();
    warp_mma_64<T>(acc, sPt, sGt, m0, g, t);
    store(2);
}

// ldmatrix variant (default).  Same five products, but the operands stay in shared memory in ONE orientation
// (q, k, v, dO as [token][d]; P and dS as [query][key], 16-bit) and every fragment is fetched with ldmatrix.x4,
// transposed by the instruction where the product needs the other orientation:
//   S  = Q K^T      A: sQ            B: sK  ([n][k] as stored)
//   dP = dO V^T     A: sG            B: sV  ([n][k] as stored)
//   dQ = dS K       A: sS            B: sK  .trans  ([k][n] as stored)
//   dK = dS^T Q     A: sS .trans     B: sQ  .trans
//   dV = P^T dO     A: sP .trans     B: sG  .trans
// No transposed copies: 55 KB of shared memory instead of 101 KB (4 CTAs per SM instead of 2) and none of the
// bank-conflicting 16-bit scatter stores the first tensor-core version needed to build them.  Rows are 144 B apart
// (72 elements): 16-byte aligned for ldmatrix and the eight row addresses of one 8 x 8 matrix fall in distinct
// 16-byte bank groups.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

constexpr size_t kAlSmem = 6 * 64 * kAmP * 2;

// acc[nt] (nt = 0..7: 16 x 8 output tiles) += A(m0.., :) * B for one warp, K = 64.
//   kTransA = false: A[m][k] = As[m][k]        kTransA = true: A[m][k] = As[k][m]
//   kTransB = false: B[k][n] = Bs[n][k]        kTransB = true: B[k][n] = Bs[k][n]
template <typename T, bool kTransA, bool kTransB>
__device__ __forceinline__ void warp_mma_ldsm(float (&acc)[8][4], uint32_t As, uint32_t Bs, int m0, int lane)
{
    const int l7 = lane & 7, mi = lane >> 3;  // matrix index 0..3 of the x4 load this lane addresses
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k0 = 16 * kk;
        uint32_t a[4];
        if (!kTransA) {
            // matrices: (m 0-7, k 0-7), (m 8-15, k 0-7), (m 0-7, k 8-15), (m 8-15, k 8-15)
            ldsm_x4(a, As + (uint32_t)(((m0 + l7 + (mi & 1) * 8) * kAmP + k0 + (mi >> 1) * 8) * 2));
        } else {
            // source rows are k: (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15), each transposed
            ldsm_x4_t(a, As + (uint32_t)(((k0 + l7 + (mi >> 1) * 8) * kAmP + m0 + (mi & 1) * 8) * 2));
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // two 8-column output tiles per x4 load
            const int n0 = 16 * np;
            uint32_t b[4];
            if (!kTransB) {
                // source rows are n: (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
                ldsm_x4(b, Bs + (uint32_t)(((n0 + l7 + (mi >> 1) * 8) * kAmP + k0 + (mi & 1) * 8) * 2));
            } else {
                // source rows are k: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15), each transposed
                ldsm_x4_t(b, Bs + (uint32_t)(((k0 + l7 + (mi & 1) * 8) * kAmP + n0 + (mi >> 1) * 8) * 2));
            }
            TrainMma<T>::mma(acc[2 * np], a, b[0], b[1]);
            TrainMma<T>::mma(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128, 4)
attention_backward_ldsm_kernel(const T *__restrict__ qkv, const T *__restrict__ g_out, T *__restrict__ g_qkv, int S, int H)
{
    extern __shared__ __align__(16) uint8_t ab_smem[];
    T *sQ = reinterpret_cast<T *>(ab_smem);
    T *sK = sQ + 64 * kAmP, *sV = sK + 64 * kAmP, *sG = sV + 64 * kAmP, *sP = sG + 64 * kAmP, *sS = sP + 64 * kAmP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int img = blockIdx.x / H, head = blockIdx.x % H;
    const int64_t row0 = (int64_t)img * S;
    const int ld = 3 * H * kAbD, ldo = H * kAbD;
    const float scale = 0.125f;
    pdl_wait();
    pdl_launch_dependents();

    // stage q, k, v, dO with 16-byte accesses (8 chunks of 8 elements per row); token rows >= S are zero
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, o = q;
        if (r < S) {
            const uint4 *src = reinterpret_cast<const uint4 *>(qkv + (row0 + r) * ld + head * kAbD);
            q = src[c];
            k = src[c + (H * kAbD) / 8];
            v = src[c + (2 * H * kAbD) / 8];
            o = reinterpret_cast<const uint4 *>(g_out + (row0 + r) * ldo + head * kAbD)[c];
        }
        *reinterpret_cast<uint4 *>(sQ + r * kAmP + 8 * c) = q;
        *reinterpret_cast<uint4 *>(sK + r * kAmP + 8 * c) = k;
        *reinterpret_cast<uint4 *>(sV + r * kAmP + 8 * c) = v;
        *reinterpret_cast<uint4 *>(sG + r * kAmP + 8 * c) = o;
    }
    __syncthreads();

    const uint32_t aQ = (uint32_t)__cvta_generic_to_shared(sQ), aK = (uint32_t)__cvta_generic_to_shared(sK);
    const uint32_t aV = (uint32_t)__cvta_generic_to_shared(sV), aG = (uint32_t)__cvta_generic_to_shared(sG);
    const uint32_t aP = (uint32_t)__cvta_generic_to_shared(sP), aS = (uint32_t)__cvta_generic_to_shared(sS);
    const int m0 = 16 * warp;
    {
        float sc[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f, dp[nt][e] = 0.f;
        warp_mma_ldsm<T, false, false>(sc, aQ, aK, m0, lane);
        warp_mma_ldsm<T, false, false>(dp, aG, aV, m0, lane);
        // element e of tile nt: row m0 + g + 8 * (e >> 1), column 8 nt + 2 t + (e & 1)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * nt + 2 * t + e;
                    float &x = sc[nt][2 * h + e];
                    x = col < S ? x * scale : -INFINITY;
                    mx = fmaxf(mx, x);
                }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float &x = sc[nt][2 * h + e];
                    x = __expf(x - mx);  // exp(-inf) = 0 for the padded columns
                    sum += x;
                }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = 1.f / sum;
            float delta = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    sc[nt][2 * h + e] *= inv;
                    delta += sc[nt][2 * h + e] * dp[nt][2 * h + e];
                }
            delta += __shfl_xor_sync(0xffffffffu, delta, 1);
            delta += __shfl_xor_sync(0xffffffffu, delta, 2);
            const int row = m0 + g + 8 * h;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int col = 8 * nt + 2 * t;
                const float p0 = sc[nt][2 * h], p1 = sc[nt][2 * h + 1];
                *reinterpret_cast<uint32_t *>(sP + row * kAmP + col) = pack2<T>(p0, p1);
                *reinterpret_cast<uint32_t *>(sS + row * kAmP + col) =
                    pack2<T>(p0 * (dp[nt][2 * h] - delta) * scale, p1 * (dp[nt][2 * h + 1] - delta) * scale);
            }
        }
    }
    __syncthreads();

    float acc[8][4];
    auto store = [&](int which) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = m0 + g + 8 * h;
            if (row < S) {
                T *dst = g_qkv + (row0 + row) * ld + which * H * kAbD + head * kAbD;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t *>(dst + 8 * nt + 2 * t) = pack2<T>(acc[nt][2 * h], acc[nt][2 * h + 1]);
            }
        }
    };
    auto clear = [&]() {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    };
    clear();
    warp_mma_ldsm<T, false, true>(acc, aS, aK, m0, lane);  // dQ rows m0..: sum_j dS[i][j] K[j][d]
    store(0);
    clear();
    warp_mma_ldsm<T, true, true>(acc, aS, aQ, m0, lane);   // dK rows m0..: sum_i dS[i][j] Q[i][d]
    store(1);
    clear();
    warp_mma_ldsm<T, true, true>(acc, aP, aG, m0, lane);   // dV rows m0..: sum_i P[i][j] dO[i][d]
    store(2);
}

static int g_attention_bwd_impl = 0;  // 0 = ldmatrix kernel, 1 = fp32 CUDA-core kernel, 2 = first mma.sync kernel (transposed copies)

}  // namespace ll

using namespace ll;

// ------------------------------------------------------------------------- class-token rows (last block of the tower)
// gather: row b of the outputs = row b * S of the inputs (two tensors at once); scatter: the reverse into zero-filled
// [B * S, W] tensors.  16-byte pieces; element type irrelevant (2-byte elements, W % 8 == 0).
__global__ void __launch_bounds__(256)
gather_cls2_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ x, uint4 *__restrict__ a_c, uint4 *__restrict__ x_c, int64_t B, int64_t S,
                   int W8)
{
    pdl_wait();
    pdl_launch_dependents();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= B * W8) return;
    const int64_t b = i / W8;
    const int c = (int)(i % W8);
    a_c[i] = a[b * S * W8 + c];
    x_c[i] = x[b * S * W8 + c];
}

__global__ void __launch_bounds__(256)
scatter_cls2_kernel(const uint4 *__restrict__ a_c, const uint4 *__restrict__ x_c, uint4 *__restrict__ a, uint4 *__restrict__ x, int64_t B, int64_t S,
                    int W8)
{
    pdl_wait();
    pdl_launch_dependents();
    const int64_t total = B * S * W8;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t row = i / W8;
        const int c = (int)(i % W8);
        const bool cls = row % S == 0;
        const int64_t b = row / S;
        a[i] = cls ? a_c[b * W8 + c] : make_uint4(0, 0, 0, 0);
        x[i] = cls ? x_c[b * W8 + c] : make_uint4(0, 0, 0, 0);
    }
}

#define LL_DISPATCH_16(dtype, CALL_F16, CALL_BF16, name)                                 \
    switch (dtype) {                                                                     \
        case LL_F16: CALL_F16; break;                                                    \
        case LL_BF16: CALL_BF16; break;                                                  \
        default: return fail(LL_ERR_INVALID, name ": dtype must be LL_F16 or LL_BF16");   \
    }

extern "C" int ll_transpose_pad_colsum(const void *src, void *dst, int64_t R, int C, int64_t Rpad, int dtype, float *colsum,
                                       ll_stream_t stream)
{
    LL_REQUIRE(src && dst && colsum && R > 0 && C > 0 && Rpad >= R, "transpose_pad_colsum: bad arguments");
    LL_REQUIRE((C % 2 == 0) && (Rpad % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0,
               "transpose_pad_colsum: needs even C and Rpad and 4-byte aligned tensors");
    dim3 grid((unsigned)((C + 63) / 64), (unsigned)((Rpad + 63) / 64));
    auto S = static_cast<const uint16_t *>(src);
    auto D = static_cast<uint16_t *>(dst);
    LL_DISPATCH_16(dtype, (transpose_pad_kernel<__half, true><<<grid, 256, 0, as_stream(stream)>>>(S, D, R, C, Rpad, colsum)),
                   (transpose_pad_kernel<__nv_bfloat16, true><<<grid, 256, 0, as_stream(stream)>>>(S, D, R, C, Rpad, colsum)), "transpose_pad_colsum");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_transpose_pad(const void *src, void *dst, int64_t R, int C, int64_t Rpad, ll_stream_t stream)
{
    LL_REQUIRE(src && dst && R > 0 && C > 0 && Rpad >= R, "transpose_pad: bad arguments");
    const bool fast = (C % 2 == 0) && (Rpad % 2 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0;
    if (fast) {
        dim3 grid((unsigned)((C + 63) / 64), (unsigned)((Rpad + 63) / 64));
        LL_CUDA_CHECK(launch_pdl(transpose_pad_kernel<__half, false>, grid, dim3(256), 0, as_stream(stream), static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), R, C, Rpad, (float *)nullptr));
    } else {
        dim3 grid((unsigned)((C + 31) / 32), (unsigned)((Rpad + 31) / 32));
        transpose_pad_scalar_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), R, C, Rpad);
    }
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_colsum(const void *src, int64_t R, int C, int dtype, float *out, ll_stream_t stream)
{
    LL_REQUIRE(src && out && R > 0 && C > 0, "colsum: bad arguments");
    cudaStream_t s = as_stream(stream);
    if (C % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // about four CTAs per SM, each thread walking its rows in batches of eight loads (64 rows of the CTA)
        const int64_t col_groups = (C + 255) / 256, target = (int64_t)num_sms() * 4;
        const int rows_per_cta = (int)std::min<int64_t>(1024, 64 * std::max<int64_t>(1, (R * col_groups / target + 63) / 64));
        dim3 grid((unsigned)col_groups, (unsigned)((R + rows_per_cta - 1) / rows_per_cta));
        LL_DISPATCH_16(dtype, LL_CUDA_CHECK(launch_pdl(colsum_vec_kernel<__half>, grid, dim3(256), 0, s, static_cast<const __half *>(src), R, C, rows_per_cta, out)),
                       LL_CUDA_CHECK(launch_pdl(colsum_vec_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, static_cast<const __nv_bfloat16 *>(src), R, C, rows_per_cta, out)), "colsum");
        LL_LAUNCH_CHECK();
        return LL_OK;
    }
    const int64_t rows_per_cta = 256;
    dim3 grid((unsigned)((C + 63) / 64), (unsigned)((R + rows_per_cta - 1) / rows_per_cta));
    LL_DISPATCH_16(dtype, (colsum_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half *>(src), R, C, rows_per_cta, out)),
                   (colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(src), R, C, rows_per_cta, out)), "colsum");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_embed_tokens(const void *patch, const void *cls, const void *pos, void *tok, int64_t B, int S, int W,
                               int dtype, ll_stream_t stream)
{
    LL_REQUIRE(patch && cls && pos && tok && B > 0 && S > 1 && W > 0 && W % 8 == 0, "embed_tokens: bad arguments");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(patch) | reinterpret_cast<uintptr_t>(cls) | reinterpret_cast<uintptr_t>(pos) |
                 reinterpret_cast<uintptr_t>(tok)) & 15) == 0, "embed_tokens: tensors must be 16-byte aligned");
    const int64_t n = B * S * (W / 8);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 16));
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, LL_CUDA_CHECK(launch_pdl(embed_tokens_kernel<__half>, dim3(grid), dim3(256), 0, s, static_cast<const __half *>(patch), static_cast<const __half *>(cls), static_cast<const __half *>(pos), static_cast<__half *>(tok), B, S, W)),
                   LL_CUDA_CHECK(launch_pdl(embed_tokens_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, static_cast<const __nv_bfloat16 *>(patch), static_cast<const __nv_bfloat16 *>(cls), static_cast<const __nv_bfloat16 *>(pos), static_cast<__nv_bfloat16 *>(tok), B, S, W)), "embed_tokens");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_embed_backward(const void *g_tok, void *g_patch, float *g_pos, float *g_cls, int64_t B, int S, int W, int dtype,
                                 ll_stream_t stream)
{
    LL_REQUIRE(g_tok && g_patch && g_pos && g_cls && B > 0 && S > 1 && W > 0 && W % 8 == 0, "embed_backward: bad arguments");
    LL_REQUIRE(S <= 65535, "embed_backward: too many token positions");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(g_tok) | reinterpret_cast<uintptr_t>(g_patch)) & 15) == 0, "embed_backward: tensors must be 16-byte aligned");
    // about four CTAs per SM; a CTA's eight batch lanes take at least eight images each
    const int64_t col_groups = (W + 255) / 256, target = (int64_t)num_sms() * 4;
    const int64_t splits = std::max<int64_t>(1, std::min<int64_t>((B + 63) / 64, target / std::max<int64_t>(1, col_groups * S)));
    const int batch_per_cta = (int)((B + splits - 1) / splits);
    dim3 grid((unsigned)col_groups, (unsigned)S, (unsigned)((B + batch_per_cta - 1) / batch_per_cta));
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, LL_CUDA_CHECK(launch_pdl(embed_backward_kernel<__half>, grid, dim3(256), 0, s, static_cast<const __half *>(g_tok), static_cast<__half *>(g_patch), g_pos, g_cls, B, S, W, batch_per_cta)),
                   LL_CUDA_CHECK(launch_pdl(embed_backward_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, static_cast<const __nv_bfloat16 *>(g_tok), static_cast<__nv_bfloat16 *>(g_patch), g_pos, g_cls, B, S, W, batch_per_cta)), "embed_backward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_quickgelu_forward(const void *u, void *f, int64_t n, int dtype, ll_stream_t stream)
{
    LL_REQUIRE(u && f && n > 0, "quickgelu_forward: bad arguments");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(f)) & 15) == 0, "quickgelu_forward: tensors must be 16-byte aligned");
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n / 8 + 255) / 256, (int64_t)num_sms() * 16));
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, LL_CUDA_CHECK(launch_pdl(quickgelu_forward_kernel<__half>, dim3(grid), dim3(256), 0, s, static_cast<const __half *>(u), static_cast<__half *>(f), n)),
                   LL_CUDA_CHECK(launch_pdl(quickgelu_forward_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, static_cast<const __nv_bfloat16 *>(u), static_cast<__nv_bfloat16 *>(f), n)), "quickgelu_forward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_quickgelu_backward(const void *u, const void *g_f, void *g_u, int64_t n, int dtype, ll_stream_t stream)
{
    LL_REQUIRE(u && g_f && g_u && n > 0, "quickgelu_backward: bad arguments");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(g_f) | reinterpret_cast<uintptr_t>(g_u)) & 15) == 0, "quickgelu_backward: tensors must be 16-byte aligned");
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n / 8 + 255) / 256, (int64_t)num_sms() * 16));
    cudaStream_t s = as_stream(stream);
    LL_DISPATCH_16(dtype, LL_CUDA_CHECK(launch_pdl(quickgelu_backward_kernel<__half>, dim3(grid), dim3(256), 0, s, static_cast<const __half *>(u), static_cast<const __half *>(g_f), static_cast<__half *>(g_u), n)),
                   LL_CUDA_CHECK(launch_pdl(quickgelu_backward_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, static_cast<const __nv_bfloat16 *>(u), static_cast<const __nv_bfloat16 *>(g_f), static_cast<__nv_bfloat16 *>(g_u), n)), "quickgelu_backward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

template <typename T>
static int launch_ln_bwd(const void *x, const void *g_h, const float *gamma, const void *g_res, void *g_x, float *g_gamma,
                         float *g_beta, int64_t rows, int width, cudaStream_t s)
{
    const unsigned grid = (unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)num_sms() * 2);  // two resident CTAs per SM
    auto X = static_cast<const T *>(x), G = static_cast<const T *>(g_h), R = static_cast<const T *>(g_res);
    auto O = static_cast<T *>(g_x);
    switch (width) {
        case 512: LL_CUDA_CHECK(launch_pdl(layernorm_backward_kernel<T, 2>, dim3(grid), dim3(256), 0, s, X, G, gamma, R, O, g_gamma, g_beta, rows)); break;
        case 768: LL_CUDA_CHECK(launch_pdl(layernorm_backward_kernel<T, 3>, dim3(grid), dim3(256), 0, s, X, G, gamma, R, O, g_gamma, g_beta, rows)); break;
        case 1024: LL_CUDA_CHECK(launch_pdl(layernorm_backward_kernel<T, 4>, dim3(grid), dim3(256), 0, s, X, G, gamma, R, O, g_gamma, g_beta, rows)); break;
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
    if (g_attention_bwd_impl == 0) {
        static bool attr3_set = false;
        if (!attr3_set) {
            LL_CUDA_CHECK(cudaFuncSetAttribute(attention_backward_ldsm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAlSmem));
            LL_CUDA_CHECK(cudaFuncSetAttribute(attention_backward_ldsm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAlSmem));
            attr3_set = true;
        }
        LL_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(g_out)) & 15) == 0, "attention_backward: qkv / g_out must be 16-byte aligned");
        LL_DISPATCH_16(dtype, LL_CUDA_CHECK(launch_pdl(attention_backward_ldsm_kernel<__half>, dim3(grid), dim3(128), (size_t)kAlSmem, s, static_cast<const __half *>(qkv), static_cast<const __half *>(g_out), static_cast<__half *>(g_qkv), S, H)),
                       LL_CUDA_CHECK(launch_pdl(attention_backward_ldsm_kernel<__nv_bfloat16>, dim3(grid), dim3(128), (size_t)kAlSmem, s, static_cast<const __nv_bfloat16 *>(qkv), static_cast<const __nv_bfloat16 *>(g_out), static_cast<__nv_bfloat16 *>(g_qkv), S, H)), "attention_backward");
        LL_LAUNCH_CHECK();
        return LL_OK;
    }
    if (g_attention_bwd_impl == 2) {
        static bool attr2_set = false;
        if (!attr2_set) {
            LL_CUDA_CHECK(cudaFuncSetAttribute(attention_backward_mma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAmSmem));
            LL_CUDA_CHECK(cudaFuncSetAttribute(attention_backward_mma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAmSmem));
            attr2_set = true;
        }
        LL_DISPATCH_16(dtype, (attention_backward_mma_kernel<__half><<<grid, 128, kAmSmem, s>>>(static_cast<const __half *>(qkv), static_cast<const __half *>(g_out), static_cast<__half *>(g_qkv), S, H)),
                       (attention_backward_mma_kernel<__nv_bfloat16><<<grid, 128, kAmSmem, s>>>(static_cast<const __nv_bfloat16 *>(qkv), static_cast<const __nv_bfloat16 *>(g_out), static_cast<__nv_bfloat16 *>(g_qkv), S, H)), "attention_backward");
        LL_LAUNCH_CHECK();
        return LL_OK;
    }
    LL_DISPATCH_16(dtype, (attention_backward_kernel<__half><<<grid, 128, kAbSmem, s>>>(static_cast<const __half *>(qkv), static_cast<const __half *>(g_out), static_cast<__half *>(g_qkv), S, H)),
                   (attention_backward_kernel<__nv_bfloat16><<<grid, 128, kAbSmem, s>>>(static_cast<const __nv_bfloat16 *>(qkv), static_cast<const __nv_bfloat16 *>(g_out), static_cast<__nv_bfloat16 *>(g_qkv), S, H)), "attention_backward");
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_attention_backward_set_impl(int impl)
{
    LL_REQUIRE(impl >= 0 && impl <= 2, "attention_backward_set_impl: 0 = ldmatrix mma.sync kernel, 1 = fp32 CUDA-core kernel, 2 = first mma.sync kernel");
    g_attention_bwd_impl = impl;
    return LL_OK;
}

extern "C" int ll_gather_cls_rows(const void *a, const void *x, void *a_c, void *x_c, int64_t batch, int S, int width, int dtype,
                                  ll_stream_t stream)
{
    LL_REQUIRE(a && x && a_c && x_c && batch > 0 && S > 0 && width > 0 && width % 8 == 0, "gather_cls_rows: bad arguments");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "gather_cls_rows: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a_c) | reinterpret_cast<uintptr_t>(x_c)) & 15) == 0,
               "gather_cls_rows: tensors must be 16-byte aligned");
    const int W8 = width / 8;
    const unsigned grid = (unsigned)((batch * W8 + 255) / 256);
    LL_CUDA_CHECK(launch_pdl(gather_cls2_kernel, dim3(grid), dim3(256), 0, as_stream(stream), static_cast<const uint4 *>(a), static_cast<const uint4 *>(x),
                             static_cast<uint4 *>(a_c), static_cast<uint4 *>(x_c), batch, (int64_t)S, W8));
    return LL_OK;
}

extern "C" int ll_scatter_cls_rows(const void *a_c, const void *x_c, void *a, void *x, int64_t batch, int S, int width, int dtype,
                                   ll_stream_t stream)
{
    LL_REQUIRE(a && x && a_c && x_c && batch > 0 && S > 0 && width > 0 && width % 8 == 0, "scatter_cls_rows: bad arguments");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "scatter_cls_rows: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a_c) | reinterpret_cast<uintptr_t>(x_c)) & 15) == 0,
               "scatter_cls_rows: tensors must be 16-byte aligned");
    const int W8 = width / 8;
    const int64_t total = batch * S * W8;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 16);
    LL_CUDA_CHECK(launch_pdl(scatter_cls2_kernel, dim3(grid), dim3(256), 0, as_stream(stream), static_cast<const uint4 *>(a_c), static_cast<const uint4 *>(x_c),
                             static_cast<uint4 *>(a), static_cast<uint4 *>(x), batch, (int64_t)S, W8));
    return LL_OK;
}
