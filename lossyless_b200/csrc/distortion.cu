// Distortion kernels of the rate-distortion training step (sm_100a), all fp32:
//   * contrastive logits tile  zs_norm @ zs_norm^T   (lossyless/distortions.py:319-324)
//   * InfoNCE cross-entropy over the gathered logits without materialising the masked copy
//     (lossyless/distortions.py:340-405), forward + backward (local tile only, as
//     lossyless/helpers.py:885-890 keeps only the local gradient of the all_gather)
//   * lossy-Z L1 distortion (lossyless/distortions.py:445, nn.PairwiseDistance(p=1))
// The logits stay fp32 on CUDA cores on purpose: they are divided by a temperature as low as
// 0.01 before the softmax, so 16-bit tensor-core operands would break the 1e-4 loss tolerance.
#include <cfloat>

#include "ll_common.cuh"

namespace ll {

// ---------------------------------------------------------------- normalise + logits
// F.normalize(zs, dim=1, p=2): x / max(||x||, 1e-12); one warp per row.
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const float *__restrict__ zs, int n, int d, float *__restrict__ zn, float *__restrict__ inv_norm)
{
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) {
        const float v = zs[(int64_t)r * d + j];
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
    for (int j = lane; j < d; j += 32) zn[(int64_t)r * d + j] = zs[(int64_t)r * d + j] * inv;
    if (lane == 0) inv_norm[r] = inv;
}

// logits[i, j] = <zn_i, zn_j>; 64x64 output tile per CTA, 4x4 per thread, K staged 16 at a time
constexpr int kLT = 64, kLK = 16;
__global__ void __launch_bounds__(256)
gram_kernel(const float *__restrict__ zn, int n, int d, float *__restrict__ logits, int ld)
{
    __shared__ float sa[kLK][kLT + 4], sb[kLK][kLT + 4];
    const int i0 = blockIdx.y * kLT, j0 = blockIdx.x * kLT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d; k0 += kLK) {
        for (int e = threadIdx.x; e < kLT * kLK; e += 256) {
            const int r = e / kLK, k = e % kLK;
            sa[k][r] = (i0 + r < n && k0 + k < d) ? zn[(int64_t)(i0 + r) * d + k0 + k] : 0.f;
            sb[k][r] = (j0 + r < n && k0 + k < d) ? zn[(int64_t)(j0 + r) * d + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kLK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = sa[k][ty * 4 + u], b[u] = sb[k][tx * 4 + u];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[u][w] = fmaf(a[u], b[w], acc[u][w]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + w;
            if (i < n && j < n) logits[(int64_t)i * ld + j] = acc[u][w];
        }
}

// d/dzs of L(logits = zn zn^T):  g_zn_i = sum_j (G_ij + G_ji) zn_j ;  then through the
// normalisation: g_x = (g - zn (zn . g)) * inv_norm.  One CTA per row i, thread per feature.
__global__ void __launch_bounds__(256)
gram_backward_kernel(const float *__restrict__ G, int ld_g, const float *__restrict__ zn,
                     const float *__restrict__ inv_norm, int n, int d, float *__restrict__ g_zs)
{
    extern __shared__ float gs[];  // n coefficients + 8 reduction slots
    float *red = gs + n;
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < n; j += blockDim.x) gs[j] = G[(int64_t)i * ld_g + j] + G[(int64_t)j * ld_g + i];
    __syncthreads();
    for (int t0 = 0; t0 < d; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        float g = 0.f;
        if (t < d)
            for (int j = 0; j < n; ++j) g = fmaf(gs[j], zn[(int64_t)j * d + t], g);
        const float zi = t < d ? zn[(int64_t)i * d + t] : 0.f;
        // block-wide dot(zn_i, g) -- only valid when d <= blockDim.x (checked by the launcher)
        float p = warp_sum(zi * g);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
        __syncthreads();
        float dot = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) dot += red[w];
        if (t < d) g_zs[(int64_t)i * d + t] = (g - zi * dot) * inv_norm[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- InfoNCE cross-entropy
// Row i of the gathered logits [n, n_cols]: own tile in columns [0, n) (column i = the row's
// own similarity, excluded), positive = column (i + n/2) mod n, remote tiles after.  One warp
// per row, online log-sum-exp.
__device__ __forceinline__ void lse_push(float &m, float &s, float v)
{
    if (v > m) {
        s = s * expf(m - v) + 1.f;
        m = v;
    } else {
        s += expf(v - m);
    }
}

__global__ void __launch_bounds__(256)
infonce_forward_kernel(const float *__restrict__ logits, int n, int n_cols, int ld, float pos_add, float inv_temp_host,
                       const float *__restrict__ inv_temp_dev,
                       float *__restrict__ loss, float *__restrict__ lse_out)
{
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int pos = (i + n / 2) % n;
    const float *row = logits + (int64_t)i * ld;
    const float inv_temp = inv_temp_dev != nullptr ? __ldg(inv_temp_dev) : inv_temp_host;
    float m = -FLT_MAX, s = 0.f;
    for (int j = lane; j < n_cols; j += 32) {
        if (j == i) continue;
        const float v = (row[j] + (j == pos ? pos_add : 0.f)) * inv_temp;
        lse_push(m, s, v);
    }
    // merge the 32 partial (max, sum) pairs
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mm = fmaxf(m, m2);
        s = s * expf(m - mm) + s2 * expf(m2 - mm);
        m = mm;
    }
    if (lane == 0) {
        const float lse = m + logf(s);
        lse_out[i] = lse;
        loss[i] = lse - (row[pos] + pos_add) * inv_temp;
    }
}

__global__ void __launch_bounds__(256)
infonce_backward_kernel(const float *__restrict__ logits, int n, int n_cols, int ld, float pos_add, float inv_temp_host,
                        const float *__restrict__ inv_temp_dev,
                        const float *__restrict__ lse, const float *__restrict__ g_loss,
                        float *__restrict__ g_local, int ld_g, float *__restrict__ g_inv_temp)
{
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int pos = (i + n / 2) % n;
    const float *row = logits + (int64_t)i * ld;
    const float gl = g_loss[i], l = lse[i];
    const float inv_temp = inv_temp_dev != nullptr ? __ldg(inv_temp_dev) : inv_temp_host;
    float gt = 0.f;  // d loss_i / d inv_temp = sum_j (p_j - [j == pos]) * (logit_j + add_j)
    for (int j = lane; j < n_cols; j += 32) {
        float gij = 0.f;
        if (j != i) {
            const float raw = row[j] + (j == pos ? pos_add : 0.f);
            const float p = expf(raw * inv_temp - l) - (j == pos ? 1.f : 0.f);
            gt = fmaf(p, raw, gt);
            gij = gl * p * inv_temp;
        }
        if (j < n) g_local[(int64_t)i * ld_g + j] = gij;
    }
    gt = warp_sum(gt);
    if (lane == 0 && g_inv_temp) atomicAdd(g_inv_temp, gl * gt);
}

// ---------------------------------------------------------------- lossy-Z (L1)
__global__ void __launch_bounds__(256)
lossy_z_forward_kernel(const float *__restrict__ zh, const float *__restrict__ z, int64_t rows, int C, float eps,
                       float *__restrict__ dist)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    float s = 0.f;
    for (int j = lane; j < C; j += 32) s += fabsf(zh[r * C + j] - z[r * C + j] + eps);
    s = warp_sum(s);
    if (lane == 0) dist[r] = s;
}

__global__ void __launch_bounds__(256)
lossy_z_backward_kernel(const float *__restrict__ zh, const float *__restrict__ z, int64_t rows, int C, float eps,
                        const float *__restrict__ g_dist, float *__restrict__ g_zh)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const float d = zh[i] - z[i] + eps;
    g_zh[i] = g_dist[i / C] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
}

}  // namespace ll

using namespace ll;

extern "C" int ll_contrastive_logits(const float *zs, int n, int d, float *zs_norm, float *inv_norm, float *logits,
                                     int ld, ll_stream_t stream)
{
    LL_REQUIRE(zs && zs_norm && inv_norm && logits && n > 0 && d > 0 && ld >= n, "contrastive_logits: bad arguments");
    cudaStream_t s = as_stream(stream);
    l2_normalize_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(zs, n, d, zs_norm, inv_norm);
    LL_LAUNCH_CHECK();
    dim3 grid((n + kLT - 1) / kLT, (n + kLT - 1) / kLT);
    gram_kernel<<<grid, 256, 0, s>>>(zs_norm, n, d, logits, ld);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_contrastive_logits_backward(const float *g_logits, int ld_g, const float *zs_norm,
                                              const float *inv_norm, int n, int d, float *g_zs, ll_stream_t stream)
{
    LL_REQUIRE(g_logits && zs_norm && inv_norm && g_zs && n > 0 && d > 0 && d <= 256 && ld_g >= n, "contrastive_logits_backward: bad arguments (d <= 256)");
    const size_t smem = (size_t)(n + 8) * sizeof(float);
    LL_REQUIRE(smem <= 48 * 1024, "contrastive_logits_backward: n = %d too large", n);
    gram_backward_kernel<<<n, 256, smem, as_stream(stream)>>>(g_logits, ld_g, zs_norm, inv_norm, n, d, g_zs);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_contrastive_ce_forward(const float *logits, int n, int n_cols, int ld, float pos_add, float inv_temp,
                                         float *loss, float *lse, ll_stream_t stream)
{
    LL_REQUIRE(logits && loss && lse && n > 1 && (n % 2) == 0 && n_cols >= n && ld >= n_cols, "contrastive_ce_forward: bad arguments");
    infonce_forward_kernel<<<(n * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(logits, n, n_cols, ld, pos_add, inv_temp, nullptr, loss, lse);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_contrastive_ce_forward_dev(const float *logits, int n, int n_cols, int ld, float pos_add,
                                             const float *inv_temp_dev, float *loss, float *lse, ll_stream_t stream)
{
    LL_REQUIRE(logits && inv_temp_dev && loss && lse && n > 1 && (n % 2) == 0 && n_cols >= n && ld >= n_cols, "contrastive_ce_forward_dev: bad arguments");
    infonce_forward_kernel<<<(n * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(logits, n, n_cols, ld, pos_add, 0.f, inv_temp_dev, loss, lse);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_contrastive_ce_backward(const float *logits, int n, int n_cols, int ld, float pos_add, float inv_temp,
                                          const float *lse, const float *g_loss, float *g_logits_local, int ld_g,
                                          float *g_inv_temp, ll_stream_t stream)
{
    LL_REQUIRE(logits && lse && g_loss && g_logits_local && n > 1 && (n % 2) == 0 && n_cols >= n && ld >= n_cols && ld_g >= n, "contrastive_ce_backward: bad arguments");
    infonce_backward_kernel<<<(n * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(logits, n, n_cols, ld, pos_add, inv_temp, nullptr, lse, g_loss, g_logits_local, ld_g, g_inv_temp);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_contrastive_ce_backward_dev(const float *logits, int n, int n_cols, int ld, float pos_add,
                                              const float *inv_temp_dev, const float *lse, const float *g_loss,
                                              float *g_logits_local, int ld_g, float *g_inv_temp, ll_stream_t stream)
{
    LL_REQUIRE(logits && inv_temp_dev && lse && g_loss && g_logits_local && n > 1 && (n % 2) == 0 && n_cols >= n && ld >= n_cols && ld_g >= n, "contrastive_ce_backward_dev: bad arguments");
    infonce_backward_kernel<<<(n * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(logits, n, n_cols, ld, pos_add, 0.f, inv_temp_dev, lse, g_loss, g_logits_local, ld_g, g_inv_temp);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_lossy_z_forward(const float *z_hat, const float *z, int64_t rows, int C, float eps, float *dist,
                                  ll_stream_t stream)
{
    LL_REQUIRE(z_hat && z && dist && rows >= 0 && C > 0, "lossy_z_forward: bad arguments");
    if (rows == 0) return LL_OK;
    lossy_z_forward_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(z_hat, z, rows, C, eps, dist);
    LL_LAUNCH_CHECK();
    return LL_OK;
}

extern "C" int ll_lossy_z_backward(const float *z_hat, const float *z, int64_t rows, int C, float eps,
                                   const float *g_dist, float *g_zhat, ll_stream_t stream)
{
    LL_REQUIRE(z_hat && z && g_dist && g_zhat && rows >= 0 && C > 0, "lossy_z_backward: bad arguments");
    if (rows == 0) return LL_OK;
    lossy_z_backward_kernel<<<(unsigned)((rows * C + 255) / 256), 256, 0, as_stream(stream)>>>(z_hat, z, rows, C, eps, g_dist, g_zhat);
    LL_LAUNCH_CHECK();
    return LL_OK;
}
