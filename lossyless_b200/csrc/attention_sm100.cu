// S <= 64-token, head-dim-64 attention on tcgen05 (sm_100a): persistent, warp-specialised.
//
// nn.MultiheadAttention core of CLIP's ResidualAttentionBlock (clip/model.py; reached from
// hub/compressor.py:93): out[b, :, h] = softmax(q k^T / 8) v for every (image b, head h).
//
// Legacy mma.sync runs at a small fraction of the tcgen05 rate on B200 (the mma.sync version of this
// kernel keeps the HMMA pipe ~100 % busy for 30-40 us per layer), so both contractions are tcgen05
// MMAs with M = 128 = two (image, head) pairs per work item:
//
//   S[128 x 128] = [Q_a; Q_b] . [K_a; K_b]^T     4 x (M128, N128, K16)  -- diagonal 64x64 blocks used
//   P            = softmax of each row over its own pair's <= 64 keys (fp32, thread per row),
//                  written block-diagonally (zeros elsewhere) as two K-major 128 x 64 tiles
//   O[128 x 64]  = P . [V_a; V_b]                8 x (M128, N64, K16), V consumed MN-major exactly
//                  as TMA lays it out (rows = keys, 128 B of head dim per row)
//
// Roles (224 threads, 2 CTAs per SM, each CTA loops over work items):
//   warp 4     TMA producer for Q,K: reloaded as soon as the previous S = QK^T has retired
//   warp 5     TMA producer for V: reloaded as soon as the previous O = PV has retired (separate warps so
//              the next item's Q,K never queue behind a V load that waits for the current O); both read the
//              [B, S, 3W] qkv tensor through a 3-D map, rows >= S zero-filled
//   warp 6     TMEM owner + MMA issuer; S(i+1) is issued before O(i) so the tensor core works on the
//              next item's scores while the softmax warps are busy with this one
//   warps 0-3  softmax + output: tcgen05.ld S -> P (fp16/bf16) -> smem; tcgen05.ld O -> scale by 1/sum
//              -> swizzled staging -> TMA store (rows >= S clipped by the 3-D tensor map)
#include <cmath>

#include "gemm_sm100.cuh"

namespace ll {

constexpr int kAtRows = 64;         // padded tokens per (image, head)
constexpr int kAtTile = 128 * 128;  // bytes of a 128-row x 64-element 16-bit tile
constexpr int kAtStage = 4 * 32 * 128;  // output staging: 4 warps x 32 rows x 128 B
constexpr int kAtSmem = 5 * kAtTile + kAtStage + 128 + 1024;  // Q, K, V, P0, P1, staging, barriers, alignment
constexpr uint32_t kAtTmemCols = 256;                          // S: columns [0,128), O: [128,192)
constexpr int kAtThreads = 224;

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

__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

enum AtBar { kQkFull = 0, kQkEmpty, kVFull, kVEmpty, kSFull, kSFree, kPFull, kPEmpty, kOFull, kOFree, kAtNumBars };

template <typename T>
__global__ void __launch_bounds__(kAtThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                    int num_pairs, int S, int H, float scale_log2e)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sq = smem, *sk = smem + kAtTile, *sv = smem + 2 * kAtTile, *sp = smem + 3 * kAtTile;
    uint8_t *stg = smem + 5 * kAtTile;
    uint64_t *bar = reinterpret_cast<uint64_t *>(stg + kAtStage);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + kAtNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = H * 64;
    if (tid == 0) {
        ptx::prefetch_tmap(&tmap_qkv);
        ptx::prefetch_tmap(&tmap_out);
        ptx::mbar_init(&bar[kQkFull], 1);
        ptx::mbar_init(&bar[kQkEmpty], 1);
        ptx::mbar_init(&bar[kVFull], 1);
        ptx::mbar_init(&bar[kVEmpty], 1);
        ptx::mbar_init(&bar[kSFull], 1);
        ptx::mbar_init(&bar[kSFree], 4);
        ptx::mbar_init(&bar[kPFull], 4);
        ptx::mbar_init(&bar[kPEmpty], 1);
        ptx::mbar_init(&bar[kOFull], 1);
        ptx::mbar_init(&bar[kOFree], 4);
        ptx::fence_barrier_init();
    }
    if (warp == 6) ptx::tmem_alloc<kAtTmemCols>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    pdl_launch_dependents();

    const int num_items = (num_pairs + 1) / 2;

    if (warp == 4) {
        // ------------------------------------------------------------ TMA producer: Q, K
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                // an odd tail repeats the last pair (its store is skipped)
                const int p0 = 2 * item, p1 = min(p0 + 1, num_pairs - 1);
                const int b0 = p0 / H, h0 = p0 % H, b1 = p1 / H, h1 = p1 % H;
                ptx::mbar_wait(&bar[kQkEmpty], (it & 1) ^ 1);
                ptx::mbar_expect_tx(&bar[kQkFull], 2 * kAtTile);
                ptx::tma_load_3d(sq, &tmap_qkv, &bar[kQkFull], h0 * 64, 0, b0);
                ptx::tma_load_3d(sq + kAtTile / 2, &tmap_qkv, &bar[kQkFull], h1 * 64, 0, b1);
                ptx::tma_load_3d(sk, &tmap_qkv, &bar[kQkFull], W + h0 * 64, 0, b0);
                ptx::tma_load_3d(sk + kAtTile / 2, &tmap_qkv, &bar[kQkFull], W + h1 * 64, 0, b1);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------------ TMA producer: V
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int p0 = 2 * item, p1 = min(p0 + 1, num_pairs - 1);
                const int b0 = p0 / H, h0 = p0 % H, b1 = p1 / H, h1 = p1 % H;
                ptx::mbar_wait(&bar[kVEmpty], (it & 1) ^ 1);
                ptx::mbar_expect_tx(&bar[kVFull], kAtTile);
                ptx::tma_load_3d(sv, &tmap_qkv, &bar[kVFull], 2 * W + h0 * 64, 0, b0);
                ptx::tma_load_3d(sv + kAtTile / 2, &tmap_qkv, &bar[kVFull], 2 * W + h1 * 64, 0, b1);
            }
        }
        __syncwarp();
    } else if (warp == 6) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr int kBf16 = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
            constexpr uint32_t idesc_s = ptx::make_idesc_f16(128, 128, kBf16);
            constexpr uint32_t idesc_o = ptx::make_idesc_f16(128, 64, kBf16, /*b_mn_major=*/1);
            const uint64_t dq = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sq));
            const uint64_t dk = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sk));
            // V read MN-major: same descriptor fields (8-row groups 1024 B apart); one K = 16 step = 16 key rows = 2048 B
            const uint64_t dv = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sv));
            const uint64_t dp0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sp));
            const uint64_t dp1 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sp + kAtTile));
            const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
            // Event-driven issue: S(i+1) = Q K^T as soon as its Q,K have landed and the softmax warps hold
            // S(i) in registers; O(i) = P V as soon as P(i), V(i) are ready and O(i-1) has been read.  Polling
            // both conditions keeps a late Q,K load for the next item from delaying this item's O.
            int next_s = 0, next_o = 0;
            while (next_o < my_items) {
                bool progressed = false;
                if (next_s < my_items && next_s <= next_o + 1) {
                    const uint32_t ph = next_s & 1;
                    if (ptx::mbar_test(&bar[kQkFull], ph) && (next_s == 0 || ptx::mbar_test(&bar[kSFree], ph ^ 1))) {
                        ptx::tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_base, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                        ptx::mma_commit(&bar[kSFull]);
                        ptx::mma_commit(&bar[kQkEmpty]);
                        ++next_s;
                        progressed = true;
                    }
                }
                if (next_o < next_s) {
                    const uint32_t ph = next_o & 1;
                    if (ptx::mbar_test(&bar[kPFull], ph) && ptx::mbar_test(&bar[kVFull], ph) && ptx::mbar_test(&bar[kOFree], ph ^ 1)) {
                        ptx::tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            ptx::mma_f16_ss(tmem_base + 128, (k < 4 ? dp0 : dp1) + 2 * (k & 3), dv + (uint64_t)(k * 128), idesc_o, k != 0);
                        ptx::mma_commit(&bar[kOFull]);
                        ptx::mma_commit(&bar[kVEmpty]);
                        ptx::mma_commit(&bar[kPEmpty]);
                        ++next_o;
                        progressed = true;
                    }
                }
                if (!progressed) __nanosleep(20);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ softmax + output (warps 0..3)
        const int half = tid >> 6;    // which pair of the item this S row belongs to
        const int row_in = tid & 63;  // token index inside the pair
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int swz = tid & 7;
        uint8_t *p_own = sp + half * kAtTile + tid * 128;
        uint8_t *p_oth = sp + (half ^ 1) * kAtTile + tid * 128;
        uint8_t *my_stg = stg + warp * (32 * 128);
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int p0 = 2 * item, p1 = min(p0 + 1, num_pairs - 1);
            const int pme = half ? p1 : p0;
            const uint32_t ph = it & 1;
            uint32_t r0[32], r1[32];
            ptx::mbar_wait(&bar[kSFull], ph);
            ptx::tc_fence_after();
            ptx::tmem_ld_32x32(taddr + half * 64, r0);
            ptx::tmem_ld_32x32(taddr + half * 64 + 32, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bar[kSFree]);

            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float v = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
                if (j < S) mx = fmaxf(mx, v);
            }
            const float neg_mx = -mx * scale_log2e;
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
            const float inv_sum = row_in < S ? 1.f / sum : 0.f;

            // P row: this pair's keys in tile `half`, zeros in the other tile (block-diagonal P)
            ptx::mbar_wait(&bar[kPEmpty], ph ^ 1);  // O(it-1) no longer reads P
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                *reinterpret_cast<uint4 *>(p_own + ((c ^ swz) << 4)) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
                *reinterpret_cast<uint4 *>(p_oth + ((c ^ swz) << 4)) = make_uint4(0, 0, 0, 0);
            }
            ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bar[kPFull]);

            ptx::mbar_wait(&bar[kOFull], ph);
            ptx::tc_fence_after();
            ptx::tmem_ld_32x32(taddr + 128, r0);
            ptx::tmem_ld_32x32(taddr + 128 + 32, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&bar[kOFree]);
                ptx::tma_store_wait_read<0>();  // this warp's previous store has drained its staging rows
            }
            __syncwarp();
            uint8_t *orow = my_stg + lane * 128;
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
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && (half == 0 || p1 != p0)) {
                // rows [32*(warp&1), +32) of pair `pme`; the map clips rows >= S
                ptx::tma_store_3d(&tmap_out, my_stg, (pme % H) * 64, (warp & 1) * 32, pme / H);
                ptx::tma_store_commit();
            }
        }
        if (lane == 0) ptx::tma_store_wait_all();
        __syncwarp();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 6) {
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
    rc = make_tmap_3d(&to, out, dtype, B, S, (int64_t)H * 64, 32);  // per-warp stores of 32 token rows
    if (rc != LL_OK) return rc;
    const int num_pairs = (int)(B * H);
    const int items = (num_pairs + 1) / 2;
    const int grid = items < 2 * num_sms() ? items : 2 * num_sms();
    const float scale_log2e = 1.4426950408889634f / sqrtf(64.f);
    LL_CUDA_CHECK(launch_pdl(kern, dim3((unsigned)grid), dim3(kAtThreads), (size_t)kAtSmem, stream, tq, to, num_pairs, S, H, scale_log2e));
    return LL_OK;
}

int launch_attention_tc(const void *qkv, void *out, int64_t B, int S, int H, int dtype, cudaStream_t stream)
{
    return dtype == LL_F16 ? launch_attention_tc_t<__half>(qkv, out, B, S, H, dtype, stream)
                           : launch_attention_tc_t<__nv_bfloat16>(qkv, out, B, S, H, dtype, stream);
}

}  // namespace ll
