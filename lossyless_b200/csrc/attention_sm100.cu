// S <= 64-token, head-dim-64 attention on tcgen05 (sm_100a).
//
// nn.MultiheadAttention core of CLIP's ResidualAttentionBlock (clip/model.py; reached from
// hub/compressor.py:93): out[b, :, h] = softmax(q k^T / 8) v for every (image b, head h).
//
// The legacy mma.sync path runs at a small fraction of the tcgen05 rate on B200 (the first
// version of this kernel kept the HMMA pipe ~100 % busy for 30-40 us per layer), so the two
// contractions are issued as tcgen05 MMAs with M = 128 = two (image, head) pairs per CTA:
//
//   S[128 x 128] = [Q_a; Q_b] . [K_a; K_b]^T     4 x (M128, N128, K16)  -- diagonal 64x64 blocks used
//   P            = softmax of each row over its own pair's <= 64 keys (fp32, thread per row),
//                  written block-diagonally (zeros elsewhere) as two K-major 128 x 64 tiles
//   O[128 x 64]  = P . [V_a; V_b]                8 x (M128, N64, K16), V consumed MN-major exactly
//                  as TMA lays it out (rows = keys, 128 B of head dim per row)
//
// Q/K/V tiles arrive by TMA from the [B, S, 3W] qkv tensor (rows >= S zero-filled by the 3-D tensor
// map), O leaves through swizzled staging + TMA store (rows >= S clipped).  Half of the S / O MMA work
// is wasted on the off-diagonal blocks; it is still several times cheaper than the HMMA path.
#include <cmath>

#include "gemm_sm100.cuh"

namespace ll {

constexpr int kAtRows = 64;          // padded tokens per (image, head)
constexpr int kAtTile = 128 * 128;   // bytes of a 128-row x 64-element 16-bit tile
constexpr int kAtSmem = 5 * kAtTile + 64 + 1024;  // Q, K, V, P0, P1 + barriers + alignment slack
constexpr uint32_t kAtTmemCols = 256;              // S: columns [0,128), O: [128,192)

template <typename T>
struct AtPack;
template <>
struct AtPack<__half> {
    static __device__ __forceinline__ uint32_t pack2(float lo, float hi)
    {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&h);
    }
};
template <>
struct AtPack<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack2(float lo, float hi)
    {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t *>(&h);
    }
};

template <typename T>
__global__ void __launch_bounds__(128)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                    int num_pairs, int S, int H, float scale_log2e)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sq = smem, *sk = smem + kAtTile, *sv = smem + 2 * kAtTile, *sp = smem + 3 * kAtTile;
    uint64_t *load_bar = reinterpret_cast<uint64_t *>(smem + 5 * kAtTile);
    uint64_t *mma_bar = load_bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = H * 64;
    if (tid == 0) {
        ptx::prefetch_tmap(&tmap_qkv);
        ptx::prefetch_tmap(&tmap_out);
        ptx::mbar_init(load_bar, 1);
        ptx::mbar_init(mma_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) ptx::tmem_alloc<kAtTmemCols>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    // the CTA's two (image, head) pairs; an odd tail repeats the last pair (its stores are harmless duplicates)
    const int p0 = 2 * blockIdx.x;
    const int p1 = min(p0 + 1, num_pairs - 1);
    const int b0 = p0 / H, h0 = p0 % H, b1 = p1 / H, h1 = p1 % H;

    if (tid == 0) {
        ptx::mbar_expect_tx(load_bar, 6 * (kAtTile / 2));
        ptx::tma_load_3d(sq, &tmap_qkv, load_bar, h0 * 64, 0, b0);
        ptx::tma_load_3d(sq + kAtTile / 2, &tmap_qkv, load_bar, h1 * 64, 0, b1);
        ptx::tma_load_3d(sk, &tmap_qkv, load_bar, W + h0 * 64, 0, b0);
        ptx::tma_load_3d(sk + kAtTile / 2, &tmap_qkv, load_bar, W + h1 * 64, 0, b1);
        ptx::tma_load_3d(sv, &tmap_qkv, load_bar, 2 * W + h0 * 64, 0, b0);
        ptx::tma_load_3d(sv + kAtTile / 2, &tmap_qkv, load_bar, 2 * W + h1 * 64, 0, b1);
        ptx::mbar_wait(load_bar, 0);
        ptx::tc_fence_after();
        // S = [Q_a; Q_b] [K_a; K_b]^T : both operands K-major, 64 = 4 x 16 along the head dim
        constexpr uint32_t idesc_s = ptx::make_idesc_f16(128, 128, std::is_same<T, __nv_bfloat16>::value ? 1 : 0);
        const uint64_t dq = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sq));
        const uint64_t dk = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sk));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_base, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        ptx::mma_commit(mma_bar);
    }
    __syncwarp();
    ptx::mbar_wait(mma_bar, 0);
    ptx::tc_fence_after();

    // ---- softmax: thread = S row; rows [0,64) belong to pair a (keys in columns [0,64)), rows [64,128) to pair b
    const int half = tid >> 6;          // which pair this row belongs to
    const int row_in = tid & 63;        // token index inside the pair
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float inv_sum = 0.f;
    {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(taddr + half * 64, r0);
        ptx::tmem_ld_32x32(taddr + half * 64 + 32, r1);
        ptx::tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const float v = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
            if (j < S) mx = fmaxf(mx, v);
        }
        float sum = 0.f;
        uint32_t packed[32];
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
            const float v0 = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
            const float v1 = __uint_as_float(j + 1 < 32 ? r0[j + 1] : r1[j + 1 - 32]);
            const float e0 = j < S ? exp2f((v0 - mx) * scale_log2e) : 0.f;
            const float e1 = j + 1 < S ? exp2f((v1 - mx) * scale_log2e) : 0.f;
            sum += e0 + e1;
            packed[j / 2] = AtPack<T>::pack2(e0, e1);
        }
        inv_sum = row_in < S ? 1.f / sum : 0.f;
        // P row: this pair's keys in tile `half`, zeros in the other tile (block-diagonal P)
        uint8_t *own = sp + half * kAtTile + tid * 128;
        uint8_t *oth = sp + (half ^ 1) * kAtTile + tid * 128;
        const int swz = tid & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<uint4 *>(own + ((c ^ swz) << 4)) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
            *reinterpret_cast<uint4 *>(oth + ((c ^ swz) << 4)) = make_uint4(0, 0, 0, 0);
        }
    }
    ptx::fence_proxy_async_smem();  // P (generic-proxy writes) -> visible to the tensor core's async proxy
    ptx::tc_fence_before();
    __syncthreads();

    if (tid == 0) {
        ptx::tc_fence_after();
        // O = P V : A = P (K-major, two 64-key tiles), B = V read MN-major (rows = keys, 128 B of head dim each):
        // one K = 16 step is 16 key rows = 2048 B of V
        constexpr uint32_t idesc_o = ptx::make_idesc_f16(128, 64, std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 1);
        const uint64_t dv = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sv));  // same LBO/SBO fields: 8-row groups 1024 B apart
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t dp = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sp + (k >> 2) * kAtTile)) + 2 * (k & 3);
            ptx::mma_f16_ss(tmem_base + 128, dp, dv + (uint64_t)(k * 128), idesc_o, k != 0);
        }
        ptx::mma_commit(mma_bar);
    }
    __syncwarp();
    ptx::mbar_wait(mma_bar, 1);
    ptx::tc_fence_after();

    // ---- O: scale by 1/sum, park in the (consumed) Q tile, TMA store per pair (rows >= S clipped by the map)
    {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(taddr + 128, r0);
        ptx::tmem_ld_32x32(taddr + 128 + 32, r1);
        ptx::tmem_ld_wait();
        uint8_t *orow = sq + tid * 128;
        const int swz = tid & 7;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = c * 8 + e * 2;
                const float v0 = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]) * inv_sum;
                const float v1 = __uint_as_float(j + 1 < 32 ? r0[j + 1] : r1[j + 1 - 32]) * inv_sum;
                w[e] = AtPack<T>::pack2(v0, v1);
            }
            *reinterpret_cast<uint4 *>(orow + ((c ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        ptx::tma_store_3d(&tmap_out, sq, h0 * 64, 0, b0);
        if (p1 != p0) ptx::tma_store_3d(&tmap_out, sq + kAtTile / 2, h1 * 64, 0, b1);
        ptx::tma_store_commit();
        ptx::tma_store_wait_all();
    }
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kAtTmemCols>(tmem_base);
    }
}

template <typename T>
static int launch_attention_tc_t(const void *qkv, void *out, int64_t B, int S, int H, int dtype, cudaStream_t stream)
{
    static bool configured = false;
    auto kern = attention_tc_kernel<T>;
    if (!configured) {
        LL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
        configured = true;
    }
    CUtensorMap tq, to;
    int rc = make_tmap_3d(&tq, qkv, dtype, B, S, 3LL * H * 64, kAtRows);
    if (rc != LL_OK) return rc;
    rc = make_tmap_3d(&to, out, dtype, B, S, (int64_t)H * 64, kAtRows);
    if (rc != LL_OK) return rc;
    const int num_pairs = (int)(B * H);
    const float scale_log2e = 1.4426950408889634f / sqrtf(64.f);
    LL_CUDA_CHECK(launch_pdl(kern, dim3((unsigned)((num_pairs + 1) / 2)), dim3(128), (size_t)kAtSmem, stream, tq, to, num_pairs, S, H,
                             scale_log2e));
    return LL_OK;
}

int launch_attention_tc(const void *qkv, void *out, int64_t B, int S, int H, int dtype, cudaStream_t stream)
{
    return dtype == LL_F16 ? launch_attention_tc_t<__half>(qkv, out, B, S, H, dtype, stream)
                           : launch_attention_tc_t<__nv_bfloat16>(qkv, out, B, S, H, dtype, stream);
}

}  // namespace ll
