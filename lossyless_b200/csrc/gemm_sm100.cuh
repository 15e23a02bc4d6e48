// tcgen05 / TMA / mbarrier PTX wrappers and the GEMM launcher interface (sm_100a).
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the driver entry point is resolved at run time)

#include "ll_common.cuh"

namespace ll {

// D[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual).  A, W, residual, D in `dtype`
// (LL_F16 / LL_BF16); bias float32 or null; act 0 none, 1 QuickGELU.
// Requirements: K % 64 == 0, N % 128 == 0, all pointers 16-byte aligned.
int gemm_launch(const void *A, const void *W, const float *bias, const void *residual, void *D, int64_t M,
                int N, int K, int dtype, int act, cudaStream_t stream);

// Same with the fused-LayerNorm epilogue operands.  stats_in: float2 [stats_in_parts][M] partial (sum, sum of squares)
// of the A rows over stat_width columns -- each accumulator row is scaled by 1/std before the bias.  stats_out:
// float2 [N / 64][M], the same partials of the D rows as stored, one part per 64 columns.  Either may be null.
int gemm_launch_ln(const void *A, const void *W, const float *bias, const void *residual, void *D, int64_t M, int N, int K,
                   int dtype, int act, const void *stats_in, int stats_in_parts, void *stats_out, int stat_width,
                   cudaStream_t stream);

// Encodes a 2-D K-major tensor map: tensor [rows, K] row-major, box [box_rows, 64], 128B swizzle.
int make_tmap_2d(CUtensorMap *map, const void *base, int dtype, int64_t rows, int64_t K, int box_rows);
// 3-D map over [batch, rows, cols] (cols contiguous), box [1, box_rows, 64], 128B swizzle: rows beyond
// `rows` of one batch entry are zero-filled on load and clipped on store (per-image token tiles).
int make_tmap_3d(CUtensorMap *map, const void *base, int dtype, int64_t batch, int64_t rows, int64_t cols, int box_rows);

// Patch embedding without im2col: D[n * 64, N] = patches(x) @ W[N, 3 P^2]^T with the A operand TMA-loaded straight from
// the image batch x [n, 3, R, R] (16-bit, `dtype`) through a 5-D tensor map.  P must be 32 (a 64-wide K block = 2 pixel
// rows of one channel of a patch, each a 64-byte SWIZZLE_64B sub-block) and the patch grid g = R / P at most 8: every
// image occupies 64 rows of D, row gy * 8 + gx for patch (gy, gx); rows with gy >= g or gx >= g are zero (TMA
// out-of-bounds fill).
int gemm_patch_launch(const void *x, const void *W, void *D, int64_t n_images, int N, int res, int patch, int dtype, cudaStream_t stream);
constexpr int kPatchRowsPerImage = 64;
constexpr int kPatchGridPitch = 8;

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LL_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LL_DONE_%=;\n\t"
        "bra LL_WAIT_%=;\n\t"
        "LL_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// non-blocking probe: true once the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled TMA load, global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// 5-D tiled TMA load (the patch-embedding operand read straight from the NCHW image batch; SASS: UTMALDG.5D)
__device__ __forceinline__ void tma_load_5d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void *smem_dst, const CUtensorMap *m, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                 int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// 2-D tiled TMA store, shared -> global, bulk-group completion (SASS: UTMASTG)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N bulk groups of this thread still have to READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before it reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- thread-block clusters / CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer) in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// Arrivals that hand the TMEM accumulator back to the MMA thread.  RELAXED on purpose: the only thing they order are
// this warp's tcgen05.ld reads, which tcgen05.wait::ld has already completed (and tcgen05.fence::before_thread_sync
// orders against the arrive); nothing in memory is published.  The .release.cluster form compiled to MEMBAR.ALL.GPU +
// ERRBAR in front of every arrive and, with the bias loads of the column group in flight, stalled each epilogue warp
// for hundreds of cycles per group (ncu source view: 20 % of all warp samples of the in-projection GEMM).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// CTA-pair TMA load: data lands in this CTA's shared memory, the transaction bytes are signalled on
// the mbarrier at `bar_cluster_addr` (the pair leader's)  (SASS: UTMALDG with .2CTA)
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *m, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *slot_in_smem)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// one MMA over the CTA pair: M = 256 (128 rows per CTA), each CTA supplies half of B's N rows
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of the pair's MMAs signalled on the same-offset mbarrier of both CTAs
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// ---- TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 in, fp32 accumulate)  (SASS: UTCHMMA)
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns  (SASS: LDTM)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, rows of 64 x 16-bit (128 B), SWIZZLE_128B as written by
// TMA; 8-row groups 1024 B apart.  (cute::UMMA::SmemDescriptor, sm_100 version bit set.)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

// K-major operand tile with rows of 32 x 16-bit (64 B), SWIZZLE_64B as written by TMA; 8-row groups 512 B apart.  Used for
// the patch-embedding A operand, whose contiguous run in memory is one 32-pixel patch row (a TMA box narrower than the
// swizzle span would be PADDED to the span in shared memory, so a 64 B run cannot be loaded into a 128 B swizzle row).
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512u >> 4) << 32;              // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)4 << 61;                        // SWIZZLE_64B
    return d;
}

// MN-major operand tile in shared memory: the tile's non-reduction dimension (M of A / N of B) is the
// contiguous one.  Built from TMA boxes of 64 (MN, contiguous = one 128 B swizzle row) x 64 (K rows), 8192 B
// each, placed one after the other along MN.  Canonical layout (cute make_umma_desc<Major::MN>, SW128), in
// 16-byte units: ((8, n), (8, k)) : ((1, LBO), (8, SBO)) -- K rows 128 B apart, 8-row K groups SBO = 1024 B
// apart, 64-element MN atoms LBO = 8192 B apart.  One K = 16 instruction step = 16 rows = 2048 B (+128 in the
// start-address field).
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)(8192u >> 4) << 16;             // leading byte offset: next 64-element MN atom
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset: next 8-row K group
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: fp32 accumulate, A K-major; B K-major or MN-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, int is_bf16, int b_mn_major = 0)
{
    return (1u << 4) | ((uint32_t)is_bf16 << 7) | ((uint32_t)is_bf16 << 10) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace ptx
}  // namespace ll
