// Persistent, warp-specialised tcgen05 GEMM for the CLIP ViT linear layers (sm_100a).
//
//   D[M,N] = act(A[M,K] @ W[N,K]^T + bias[N]) (+ residual[M,N])
//
// Every dense contraction of the reference's encoder -- patch-embed conv (as im2col GEMM),
// attention in/out projections, MLP c_fc / c_proj and the final `@ proj`
// (clip/model.py VisionTransformer / ResidualAttentionBlock, reached from
// hub/compressor.py:93 and lossyless/architectures.py:351-354) -- has this shape with both
// operands K-major, which is what nn.Linear stores.
//
// Roles (192 threads, 1 CTA per SM, grid = min(#tiles, #SMs), static round-robin tiles):
//   warp 4      TMA producer: 128xBK A tile + BNxBK W tile per stage, 128B swizzle, mbarrier tx
//   warp 5      allocates TMEM; one lane issues tcgen05.mma (M=128, N=BN, K=16) into one of two
//               TMEM accumulator buffers, commits stage-empty / accumulator-full barriers
//   warps 0..3  epilogue, overlapping the next tile's main loop (double-buffered TMEM): per
//               32-row x 64-column group tcgen05.ld -> bias / QuickGELU / residual in fp32 ->
//               16-bit, swizzled st.shared -> TMA store (full 128 B lines; the row-per-thread
//               accumulator layout would otherwise scatter 16 B pieces over 32 lines per store).
//               The residual tile arrives by TMA into the same staging buffer one group ahead.
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "gemm_sm100.cuh"

namespace ll {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 x 16-bit = one 128 B swizzle row
constexpr int kEpiWarps = 8;  // two epilogue warps per TMEM lane quarter, alternating 64-column groups
constexpr int kGemmThreads = (kEpiWarps + 2) * 32;
// Warp roles.  The SM sub-partition schedulers favour the higher warp id among eligible warps, and
// warp w runs on sub-partition w % 4: the two latency-critical single-thread roles get the HIGHEST
// ids so the ALU-heavy epilogue warps sharing their sub-partitions (0 and 1) cannot starve them.
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kABytes = kBM * kBK * 2;
constexpr int kEpiCols = 64;                    // one 128 B swizzle row of 16-bit outputs
constexpr int kEpiBufBytes = 32 * kEpiCols * 2;  // 32 rows (one warp's TMEM lanes) x 64 columns

// kCtas = 1: one CTA per 128 x BN tile.  kCtas = 2: a CTA pair (cluster of 2, tcgen05
// cta_group::2) per 256 x BN tile -- each CTA holds its own 128 A rows and HALF of the W tile, so per
// CTA and K-block the shared-memory traffic drops from (128 + BN) to (128 + BN/2) rows while the
// tensor cores of both SMs run the same M = 256 instruction.
template <int BN, int kCtas>
struct GemmShape {
    static constexpr int kBRows = BN / kCtas;  // W rows staged per CTA
    static constexpr int kBBytes = kBRows * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kAccStride = (BN == 192) ? 256 : BN;  // TMEM columns between the two accumulators
    static constexpr int kTmemCols = 2 * kAccStride;           // power of two >= 32
    static constexpr int kGroups = BN / kEpiCols;              // 64-column epilogue groups per tile
    static constexpr int kStagingBytes = 4 * kGroups * kEpiBufBytes;  // 4 epilogue warps x one buffer per group
    static constexpr int kMaxSmem = 232448 - 1024;             // 227 KB minus the alignment slack
    static constexpr int kStagesFit = (kMaxSmem - kStagingBytes - 512) / kStageBytes;
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kBarBytes = (2 * kStages + 4 + 4 * kGroups) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;  // +1024: alignment
    static_assert(kStages >= 3, "not enough shared memory for a 3-stage pipeline");
    static_assert(kBarBytes <= 512, "barrier block overflows its reservation");
};

template <typename T>
struct Pack8;  // 8 x 16-bit <-> 8 floats
template <>
struct Pack8<__half> {
    static __device__ __forceinline__ void unpack(const uint4 &v, float (&f)[8])
    {
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 t = __half22float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[8])
    {
        uint4 v;
        __half2 *h = reinterpret_cast<__half2 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};
template <>
struct Pack8<__nv_bfloat16> {
    static __device__ __forceinline__ void unpack(const uint4 &v, float (&f)[8])
    {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 t = __bfloat1622float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[8])
    {
        uint4 v;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};

__device__ __forceinline__ float quick_gelu(float x)
{
    // CLIP QuickGELU: x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x)); one MUFU op per element
    // (the epilogue of c_fc is SFU-bound otherwise: ex2 + rcp = 2 MUFU x 32 768 elements per tile)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// d/dx [x sigmoid(1.702 x)] = s + 1.702 x s (1 - s) with s = 0.5 (1 + tanh(0.851 x)): one MUFU op per element
__device__ __forceinline__ float quick_gelu_grad(float x)
{
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    const float sg = fmaf(0.5f, t, 0.5f);
    return fmaf(1.702f * x * sg, 1.f - sg, sg);
}

// QuickGELU on two fp16 values at once, AFTER the rounding to the output type: 2 HMUL2 + 1 MUFU (tanh.approx.f16x2)
// + 1 HFMA2 per PAIR instead of 3 FP32 ops + 1 MUFU per ELEMENT.  The c_fc epilogue (K = 768: 3.3 us of MMA per
// tile) was the one epilogue that did not hide behind the tensor pipe (ncu: 66 % tensor-active vs 75-79 % for the
// in-projection with the same K).  The inputs are converted with .satfinite so that a -inf cannot appear
// (0.5 x (1 + tanh) would be inf - inf); the result differs from the fp32 evaluation by <= 2 fp16 ulp.
__device__ __forceinline__ uint32_t cvt_f16x2_satfinite(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t quick_gelu_h2(uint32_t x)
{
    const __half2 xv = *reinterpret_cast<const __half2 *>(&x);
    const __half2 a = __hmul2(xv, __float2half2_rn(0.851f));
    uint32_t t;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t *>(&a)));
    const __half2 hx = __hmul2(xv, __float2half2_rn(0.5f));
    const __half2 o = __hfma2(hx, *reinterpret_cast<const __half2 *>(&t), hx);
    return *reinterpret_cast<const uint32_t *>(&o);
}

template <typename T, int BN, int kCtas, bool kDual = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r,
                    const float *__restrict__ bias, int has_residual, int M, int N, int K, int act,
                    const float2 *__restrict__ stats_in, int stats_in_parts, float2 *__restrict__ stats_out,
                    float inv_stat_width)
{
    using S = GemmShape<BN, kCtas>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sm_a = smem;
    uint8_t *sm_b = smem + S::kStages * kABytes;
    uint8_t *sm_epi = smem + S::kStages * S::kStageBytes;  // [4 warps][2][32 rows x 128 B], 1024-aligned
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(sm_epi + S::kStagingBytes);
    uint64_t *empty_bar = full_bar + S::kStages;
    uint64_t *tfull_bar = empty_bar + S::kStages;  // [2] accumulator ready for the epilogue
    uint64_t *tempty_bar = tfull_bar + 2;          // [2] accumulator drained
    uint64_t *res_bar = tempty_bar + 2;            // [4 warps][kGroups] residual tile landed in the staging buffer
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_bar + 4 * S::kGroups);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = kCtas == 2 ? ptx::cluster_ctarank() : 0;  // 0 = pair leader (issues the MMAs)
    const int worker = blockIdx.x / kCtas;       // persistent worker (CTA or CTA pair)
    const int num_workers = gridDim.x / kCtas;

    if (warp == kProducerWarp && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w);
        ptx::prefetch_tmap(&tmap_d);
        if (has_residual) ptx::prefetch_tmap(&tmap_r);
        for (int i = 0; i < 4 * S::kGroups; ++i) ptx::mbar_init(&res_bar[i], 1);
        for (int i = 0; i < S::kStages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            // one arrive per epilogue warp that owns at least one column group (of both CTAs)
            ptx::mbar_init(&tempty_bar[i], (S::kGroups >= 2 ? kEpiWarps : 4) * kCtas);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        if constexpr (kCtas == 2)
            ptx::tmem_alloc_pair<S::kTmemCols>(tmem_slot);
        else
            ptx::tmem_alloc<S::kTmemCols>(tmem_slot);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kCtas == 2) ptx::cluster_sync();  // peer's barriers are initialised, its TMEM allocated
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous
    // kernel's tail; from here on global memory written by predecessors is touched
    pdl_wait();
    pdl_launch_dependents();

    constexpr int kTileM = kBM * kCtas;
    const int num_n = N / BN;
    const int num_m = (M + kTileM - 1) / kTileM;
    const int num_tiles = num_m * num_n;
    const int num_kb = K / kBK;
    const bool tn = (act & 0x2000) != 0;  // both operands MN-major (transposed): D = A^T B
    // patch-embedding mode: A is the image batch itself (5-D map); bits 16.. of `act` hold the K blocks per channel
    const bool patch_mode = (act & 0x8000) != 0;
    const int kb_per_channel = act >> 16;
    // B alone MN-major (act bit 0x40000): D = A B with B given as [K, N] row-major -- the data gradient dX = dY W read from
    // the weight as nn.Linear stores it ([out, in] = [K, N]), no transposed copy
    const bool bmn = (act & 0x40000) != 0 && !tn && !patch_mode;
    // dual output (act bit 0x80000, c_fc of the fine-tuning forward): tmap_d receives the pre-activation u = A W^T + b,
    // tmap_r (free: no residual in this mode) the activation QuickGELU(u) -- both are needed by the backward pass
    // (a template parameter: its 32 extra live registers in the epilogue must not cost the other instantiations anything)
    constexpr bool dual = kDual;
    // split-K (act bit 0x10000; weight gradients: few output tiles, a very deep reduction): work item = (tile, split),
    // split s reduces K blocks [s nkb / S, (s + 1) nkb / S) and leaves a float32 partial tile in the workspace
    // (passed through the stats_out pointer, S through stats_in_parts); a second kernel adds the partials
    const bool splitk = (act & 0x10000) != 0 && !patch_mode;
    const int splits = splitk ? stats_in_parts : 1;
    const int num_work = num_tiles * splits;

    if (warp == kProducerWarp) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int work = worker; work < num_work; work += num_workers) {
                const int tile = work % num_tiles, sp = work / num_tiles;
                const int m0 = (tile / num_n) * kTileM + (int)cta_rank * kBM;
                const int n0 = (tile % num_n) * BN + (int)cta_rank * S::kBRows;
                const int kb_begin = (int)((int64_t)sp * num_kb / splits), kb_end = (int)((int64_t)(sp + 1) * num_kb / splits);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (tn) {
                        // TN operands (D = A^T B, weight gradients): A is [K rows, M columns], B is [K rows, N columns];
                        // the same stage bytes arrive as 64-column x 64-row boxes, one per 64-element MN atom
                        // (rows past the end of K are zero-filled by TMA, so K needs no padding)
                        if constexpr (kCtas == 2) {
                            const uint32_t lead_bar = ptx::map_to_cta(&full_bar[stage], 0);
                            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                            for (int h = 0; h < kBM / 64; ++h)
                                ptx::tma_load_2d_pair(sm_a + stage * kABytes + h * 8192, &tmap_a, lead_bar, m0 + 64 * h, kb * kBK);
                            for (int h = 0; h < S::kBRows / 64; ++h)
                                ptx::tma_load_2d_pair(sm_b + stage * S::kBBytes + h * 8192, &tmap_w, lead_bar, n0 + 64 * h, kb * kBK);
                        } else {
                            ptx::mbar_expect_tx(&full_bar[stage], S::kStageBytes);
                            for (int h = 0; h < kBM / 64; ++h)
                                ptx::tma_load_2d(sm_a + stage * kABytes + h * 8192, &tmap_a, &full_bar[stage], m0 + 64 * h, kb * kBK);
                            for (int h = 0; h < S::kBRows / 64; ++h)
                                ptx::tma_load_2d(sm_b + stage * S::kBBytes + h * 8192, &tmap_w, &full_bar[stage], n0 + 64 * h, kb * kBK);
                        }
                    } else if (bmn) {
                        // A K-major as usual; B arrives as 64-column x 64-row boxes of the [K, N] matrix, one per 64-element
                        // N atom (MN-major SWIZZLE_128B layout, as in the TN mode)
                        if constexpr (kCtas == 2) {
                            const uint32_t lead_bar = ptx::map_to_cta(&full_bar[stage], 0);
                            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                            ptx::tma_load_2d_pair(sm_a + stage * kABytes, &tmap_a, lead_bar, kb * kBK, m0);
                            for (int h = 0; h < S::kBRows / 64; ++h)
                                ptx::tma_load_2d_pair(sm_b + stage * S::kBBytes + h * 8192, &tmap_w, lead_bar, n0 + 64 * h, kb * kBK);
                        } else {
                            ptx::mbar_expect_tx(&full_bar[stage], S::kStageBytes);
                            ptx::tma_load_2d(sm_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * kBK, m0);
                            for (int h = 0; h < S::kBRows / 64; ++h)
                                ptx::tma_load_2d(sm_b + stage * S::kBBytes + h * 8192, &tmap_w, &full_bar[stage], n0 + 64 * h, kb * kBK);
                        }
                    } else if (patch_mode) {
                        // im2col-free patch embedding: this CTA's 128 A rows are the 8 x 8 (padded) patch grids of two
                        // images; K block kb = pixel rows py0, py0 + 1 of channel ch of every patch.  One box (fastest
                        // first: [32 px][8 gx][1 py][8 gy][1 image-channel]) is 64 rows x 64 B, K-major, SWIZZLE_64B;
                        // the stage holds [pixel row][image][64 rows x 64 B].
                        const int ch = kb / kb_per_channel, py0 = (kb % kb_per_channel) * 2;
                        const int img = m0 / kPatchRowsPerImage;
                        uint8_t *dst = sm_a + stage * kABytes;
                        if constexpr (kCtas == 2) {
                            const uint32_t lead_bar = ptx::map_to_cta(&full_bar[stage], 0);
                            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                ptx::tma_load_5d_pair(dst + q * 4096, &tmap_a, lead_bar, 0, 0, py0 + (q >> 1), 0, 3 * (img + (q & 1)) + ch);
                            ptx::tma_load_2d_pair(sm_b + stage * S::kBBytes, &tmap_w, lead_bar, kb * kBK, n0);
                        } else {
                            ptx::mbar_expect_tx(&full_bar[stage], S::kStageBytes);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                ptx::tma_load_5d(dst + q * 4096, &tmap_a, &full_bar[stage], 0, 0, py0 + (q >> 1), 0, 3 * (img + (q & 1)) + ch);
                            ptx::tma_load_2d(sm_b + stage * S::kBBytes, &tmap_w, &full_bar[stage], kb * kBK, n0);
                        }
                    } else if constexpr (kCtas == 2) {
                        // both CTAs' bytes are counted on the LEADER's full barrier, which alone arms it
                        const uint32_t lead_bar = ptx::map_to_cta(&full_bar[stage], 0);
                        if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                        ptx::tma_load_2d_pair(sm_a + stage * kABytes, &tmap_a, lead_bar, kb * kBK, m0);
                        ptx::tma_load_2d_pair(sm_b + stage * S::kBBytes, &tmap_w, lead_bar, kb * kBK, n0);
                    } else {
                        ptx::mbar_expect_tx(&full_bar[stage], S::kStageBytes);
                        ptx::tma_load_2d(sm_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * kBK, m0);
                        ptx::tma_load_2d(sm_b + stage * S::kBBytes, &tmap_w, &full_bar[stage], kb * kBK, n0);
                    }
                    if (++stage == S::kStages) stage = 0, phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------ MMA issuer (pair leader only)
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(kTileM, BN, std::is_same<T, __nv_bfloat16>::value ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int work = worker; work < num_work; work += num_workers, ++it) {
                const int sp = work / num_tiles;
                const int kb_begin = (int)((int64_t)sp * num_kb / splits), kb_end = (int)((int64_t)(sp + 1) * num_kb / splits);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * S::kAccStride;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t addr_a = ptx::smem_u32(sm_a + stage * kABytes), addr_b = ptx::smem_u32(sm_b + stage * S::kBBytes);
                    const uint64_t da = tn ? ptx::make_mnmajor_sw128_desc(addr_a) : ptx::make_kmajor_sw128_desc(addr_a);
                    const uint64_t db = (tn || bmn) ? ptx::make_mnmajor_sw128_desc(addr_b) : ptx::make_kmajor_sw128_desc(addr_b);
                    // per K=16 step: K-major +32 B inside the 128 B swizzle row (start-address field += 2);
                    // MN-major 16 K rows further down (2048 B, field += 128)
                    const uint64_t kstep = tn ? 128 : 2;
                    const uint64_t kstep_b = (tn || bmn) ? 128 : 2;
                    const uint32_t id = tn ? (idesc | (1u << 15) | (1u << 16)) : bmn ? (idesc | (1u << 16)) : idesc;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        // patch mode: A is two SWIZZLE_64B sub-blocks of 32 K-columns (8192 B apart), two K = 16 steps each
                        const uint64_t dak = patch_mode ? ptx::make_kmajor_sw64_desc(addr_a + (k >> 1) * 8192) + 2 * (k & 1) : da + kstep * k;
                        if constexpr (kCtas == 2)
                            ptx::mma_f16_ss_pair(tmem_d, dak, db + kstep_b * k, id, ((kb - kb_begin) | k) != 0);
                        else
                            ptx::mma_f16_ss(tmem_d, dak, db + kstep_b * k, id, ((kb - kb_begin) | k) != 0);
                    }
                    // frees the smem stage (in both CTAs) when the MMAs retire
                    if constexpr (kCtas == 2)
                        ptx::mma_commit_pair(&empty_bar[stage]);
                    else
                        ptx::mma_commit(&empty_bar[stage]);
                    if (++stage == S::kStages) stage = 0, phase ^= 1;
                }
                if constexpr (kCtas == 2)
                    ptx::mma_commit_pair(&tfull_bar[acc]);
                else
                    ptx::mma_commit(&tfull_bar[acc]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (warps 0..3)
        // One 32-row x 64-column staging buffer per column group and warp.  With a residual, all of
        // the tile's residual groups are TMA-loaded into them as soon as the previous tile's stores
        // have drained -- i.e. while the MMAs of this tile are still running -- and the results
        // overwrite them in place before the TMA store.
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter + 32)
        const int gpar = warp >> 2;    // this warp owns the column groups g with (g & 1) == gpar
        constexpr int kMaxOwn = (S::kGroups + 1) / 2;
        const int n_own = (S::kGroups - gpar + 1) / 2;  // groups this warp owns per tile
        uint8_t *buf0 = sm_epi + quarter * S::kGroups * kEpiBufBytes;
        uint64_t *rbar = res_bar + quarter * S::kGroups;
        uint32_t rphase = 0;           // all of a tile's residual barriers flip together
        const int swz = lane & 7;      // 128 B swizzle: 16 B chunk j of row r lives at chunk j ^ (r & 7)
        uint8_t *my_row0 = buf0 + lane * 128;
        int it = 0;
        const uint32_t tempty_lead[2] = {kCtas == 2 ? ptx::map_to_cta(&tempty_bar[0], 0) : 0,
                                         kCtas == 2 ? ptx::map_to_cta(&tempty_bar[1], 0) : 0};
        for (int work = worker; n_own > 0 && work < num_work; work += num_workers, ++it) {
            const int tile = work % num_tiles, sp = work / num_tiles;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (tile / num_n) * kTileM + (int)cta_rank * kBM + quarter * 32;
            const int n0 = (tile % num_n) * BN;
            if (has_residual && lane == 0) {
                ptx::tma_store_wait_read<0>();  // this warp's previous stores no longer read its buffers
#pragma unroll
                for (int g = gpar; g < S::kGroups; g += 2) {
                    ptx::mbar_expect_tx(&rbar[g], kEpiBufBytes);
                    ptx::tma_load_2d(buf0 + g * kEpiBufBytes, &tmap_r, &rbar[g], n0 + g * kEpiCols, m0);
                }
            }
            // fused LayerNorm, consumer side: the weights were row-centred and scaled by gamma on the host, so
            // LN(x) W^T + b = rstd(x) * (x W''^T) + b' -- only the row's 1/std is needed here, from the
            // (sum, sum of squares) partials the producing kernel's epilogue left, one [M] plane per part, added in
            // a fixed order (no atomics: the encoder's bytes stay reproducible run to run)
            float row_scale = 1.f;
            if (stats_in != nullptr && !splitk) {
                const int r = m0 + lane;
                if (r < M) {
                    // all (<= 16) part loads in flight together, then a fixed-order sum
                    float2 st[16];
#pragma unroll
                    for (int p = 0; p < 16; ++p)
                        st[p] = p < stats_in_parts ? __ldg(&stats_in[(size_t)p * M + r]) : make_float2(0.f, 0.f);
                    float sx = 0.f, sq = 0.f;
#pragma unroll
                    for (int p = 0; p < 16; ++p) sx += st[p].x, sq += st[p].y;
                    const float mu = sx * inv_stat_width;
                    row_scale = rsqrtf(fmaxf(sq * inv_stat_width - mu * mu, 0.f) + 1e-5f);
                }
            }
            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * S::kAccStride;
#pragma unroll 1
            for (int g = gpar; g < S::kGroups; g += 2) {
                uint8_t *my_row = my_row0 + g * kEpiBufBytes;
                // bias for the group's 64 columns first: the (L1-resident, warp-uniform) loads complete
                // under the TMEM load latency instead of stalling the arithmetic below
                float4 bv[16];
                if (bias != nullptr && !(act & 0x400)) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) bv[q] = __ldg(reinterpret_cast<const float4 *>(bias + n0 + g * kEpiCols) + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) bv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float st_sum = 0.f, st_sq = 0.f;
                uint32_t r0[32], r1[32];
                if (!(act & 0x800)) {
                    ptx::tmem_ld_32x32(taddr + g * kEpiCols, r0);
                    ptx::tmem_ld_32x32(taddr + g * kEpiCols + 32, r1);
                }
                if (!has_residual && lane == 0) {
                    // buffer g was last read by this warp's store of group g in the previous tile: at most
                    // its (groups per tile - 1) younger stores may still be pending
                    if (n_own == kMaxOwn)
                        ptx::tma_store_wait_read<kMaxOwn - 1>();
                    else
                        ptx::tma_store_wait_read<(kMaxOwn >= 2 ? kMaxOwn - 2 : 0)>();
                }
                ptx::tmem_ld_wait();
                if (g + 2 >= S::kGroups) {
                    // accumulator fully read: hand the TMEM buffer back to the MMA warp early
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (kCtas == 2)
                            ptx::mbar_arrive_cluster(tempty_lead[acc]);  // the leader's MMA thread waits for both CTAs
                        else
                            ptx::mbar_arrive_relaxed(&tempty_bar[acc]);
                    }
                }
                if (has_residual) {
                    ptx::mbar_wait(&rbar[g], rphase);
                } else {
                    __syncwarp();  // orders lane 0's wait_group.read before the overwrite of buffer g
                }
                if (splitk) {
                    // float32 partial of this split: row m0 + lane, 64 consecutive columns, straight from the registers
                    if (m0 + lane < M) {
                        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(stats_out) + ((size_t)sp * M + (m0 + lane)) * N + n0 + g * kEpiCols);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            dst[j] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]), __uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3]));
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            dst[8 + j] = make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]), __uint_as_float(r1[4 * j + 2]), __uint_as_float(r1[4 * j + 3]));
                        }
                    }
                    continue;
                }
                uint4 gp[8];  // dual output: the activation's 8 x 16-byte pieces wait here until the u tile has left the buffer
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(j < 4 ? r0[j * 8 + e] : r1[(j - 4) * 8 + e]);
                    {
                        const float4 b0 = bv[2 * j], b1 = bv[2 * j + 1];
                        f[0] = fmaf(f[0], row_scale, b0.x), f[1] = fmaf(f[1], row_scale, b0.y);
                        f[2] = fmaf(f[2], row_scale, b0.z), f[3] = fmaf(f[3], row_scale, b0.w);
                        f[4] = fmaf(f[4], row_scale, b1.x), f[5] = fmaf(f[5], row_scale, b1.y);
                        f[6] = fmaf(f[6], row_scale, b1.z), f[7] = fmaf(f[7], row_scale, b1.w);
                    }
                    uint4 *slot = reinterpret_cast<uint4 *>(my_row + ((j ^ swz) << 4));
                    if constexpr (std::is_same<T, __half>::value) {
                        if ((act & 0xff) == 1 && !has_residual && stats_out == nullptr && !(act & 0x4000) && !dual) {
                            uint4 packed;
                            packed.x = quick_gelu_h2(cvt_f16x2_satfinite(f[0], f[1]));
                            packed.y = quick_gelu_h2(cvt_f16x2_satfinite(f[2], f[3]));
                            packed.z = quick_gelu_h2(cvt_f16x2_satfinite(f[4], f[5]));
                            packed.w = quick_gelu_h2(cvt_f16x2_satfinite(f[6], f[7]));
                            if (!(act & 0x200)) *slot = packed;
                            continue;
                        }
                    }
                    if (dual) {
                        // u first (rounded to the output type, as the separate activation pass would read it back)
                        const uint4 pre = Pack8<T>::pack(f);
                        *slot = pre;
                        float uf[8];
                        Pack8<T>::unpack(pre, uf);
#pragma unroll
                        for (int e = 0; e < 8; ++e) uf[e] = quick_gelu(uf[e]);
                        gp[j] = Pack8<T>::pack(uf);
                        continue;
                    }
                    if ((act & 0xff) == 1) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = quick_gelu(f[e]);
                    }
                    if (has_residual) {
                        float rf[8];
                        Pack8<T>::unpack(*slot, rf);
                        if ((act & 0xff) == 2) {
                            // QuickGELU backward fused into the data-gradient GEMM of c_proj: the "residual" tile is
                            // the pre-activation u, the result g_u = g_f * QuickGELU'(u)
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] *= quick_gelu_grad(rf[e]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] += rf[e];
                        }
                    }
                    const uint4 packed = Pack8<T>::pack(f);
                    if (!(act & 0x200)) *slot = packed;
                    if (stats_out != nullptr) {
                        // fused LayerNorm, producer side.  Taken from the fp32 values before the 16-bit rounding: the
                        // rounding errors average out over the row (mean moves by ~1e-5 sigma, 1/std by ~1e-7
                        // relative), far below the 16-bit rounding of the consumer's own output
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            st_sum += f[e];
                            st_sq = fmaf(f[e], f[e], st_sq);
                        }
                    }
                }
                if (dual) {
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tmap_d, buf0 + g * kEpiBufBytes, n0 + g * kEpiCols, m0);
                        ptx::tma_store_commit();
                        ptx::tma_store_wait_read<0>();  // the u tile has been read out of the staging buffer
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4 *>(my_row + ((j ^ swz) << 4)) = gp[j];
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_2d(&tmap_r, buf0 + g * kEpiBufBytes, n0 + g * kEpiCols, m0);
                        ptx::tma_store_commit();
                    }
                    continue;
                }
                if (stats_out != nullptr && m0 + lane < M) {
                    // one part per 64-column group -- a single writer per slot and a summation order that does not
                    // depend on the tile shape, so the encoder's bits are reproducible and batch-size invariant
                    stats_out[(size_t)((n0 >> 6) + g) * M + m0 + lane] = make_float2(st_sum, st_sq);
                }
                if (!(act & 0x1000)) ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && !(act & 0x100)) {
                    ptx::tma_store_2d(&tmap_d, buf0 + g * kEpiBufBytes, n0 + g * kEpiCols, m0);
                    ptx::tma_store_commit();
                }
            }
            rphase ^= 1;
        }
        if (lane == 0) ptx::tma_store_wait_all();  // global writes complete before the CTA exits
        __syncwarp();
    }

    // ---------------------------------------------------------------- teardown
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (kCtas == 2) ptx::cluster_sync();  // the peer may still be read by / signalling this CTA
    if (warp == kMmaWarp) {
        ptx::tc_fence_after();
        if constexpr (kCtas == 2)
            ptx::tmem_dealloc_pair<S::kTmemCols>(tmem_base);
        else
            ptx::tmem_dealloc<S::kTmemCols>(tmem_base);
    }
}

// --------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tensor maps are pure functions of (pointer, shape, box): the encoder re-launches the same ~60 GEMMs on
// the same workspace every pass, so the driver-side encode (~1-2 us each, 4 per GEMM) is cached.
struct TmapKey {
    const void *base;
    int64_t d0, d1, d2;
    int box_rows, dtype, rank;
    bool operator==(const TmapKey &o) const
    {
        return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && box_rows == o.box_rows && dtype == o.dtype && rank == o.rank;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey &k) const
    {
        size_t h = reinterpret_cast<size_t>(k.base);
        for (int64_t v : {k.d0, k.d1, k.d2, (int64_t)k.box_rows, (int64_t)k.dtype, (int64_t)k.rank}) h = h * 1000003u ^ (size_t)v;
        return h;
    }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

static bool tmap_cache_get(const TmapKey &k, CUtensorMap *out)
{
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    auto it = g_tmap_cache.find(k);
    if (it == g_tmap_cache.end()) return false;
    *out = it->second;
    return true;
}
static void tmap_cache_put(const TmapKey &k, const CUtensorMap &m)
{
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (g_tmap_cache.size() > 4096) g_tmap_cache.clear();
    g_tmap_cache[k] = m;
}

int make_tmap_2d(CUtensorMap *map, const void *base, int dtype, int64_t rows, int64_t K, int box_rows)
{
    const TmapKey key{base, K, rows, 1, box_rows, dtype, 2};
    if (tmap_cache_get(key, map)) return LL_OK;
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(LL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    LL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    LL_REQUIRE(K % kBK == 0 && rows > 0 && box_rows > 0 && box_rows <= 256, "make_tmap_2d: bad shape rows=%lld K=%lld box=%d", (long long)rows, (long long)K, box_rows);
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = dtype == LL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    tmap_cache_put(key, *map);
    return LL_OK;
}

int make_tmap_3d(CUtensorMap *map, const void *base, int dtype, int64_t batch, int64_t rows, int64_t cols, int box_rows)
{
    const TmapKey key{base, cols, rows, batch, box_rows, dtype, 3};
    if (tmap_cache_get(key, map)) return LL_OK;
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(LL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    LL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    LL_REQUIRE(cols % 64 == 0 && rows > 0 && batch > 0 && box_rows > 0 && box_rows <= 256, "make_tmap_3d: bad shape");
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = dtype == LL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
    tmap_cache_put(key, *map);
    return LL_OK;
}

// 5-D map over the image batch x [n, 3, R, R] for the patch-embedding A operand.  Dimensions, fastest first:
// px [P] (stride 1 element), gx [g] (stride P), py [P] (stride R), gy [g] (stride P R), image-channel [3 n] (stride R R);
// box [P, 8, 1, 8, 1] = 64 rows (gy, gx) x 64 B, SWIZZLE_64B.
static int make_tmap_patch5d(CUtensorMap *map, const void *x, int dtype, int64_t n, int res, int patch)
{
    const TmapKey key{x, res, patch, n, 0, dtype, 5};
    if (tmap_cache_get(key, map)) return LL_OK;
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(LL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    LL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    const int g = res / patch;
    const cuuint64_t dims[5] = {(cuuint64_t)patch, (cuuint64_t)g, (cuuint64_t)patch, (cuuint64_t)g, (cuuint64_t)(3 * n)};
    const cuuint64_t strides[4] = {(cuuint64_t)patch * 2, (cuuint64_t)res * 2, (cuuint64_t)patch * res * 2, (cuuint64_t)res * res * 2};
    const cuuint32_t box[5] = {(cuuint32_t)patch, (cuuint32_t)kPatchGridPitch, 1, (cuuint32_t)kPatchGridPitch, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = dtype == LL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(map, dt, 5, const_cast<void *>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ERR_CUDA, "cuTensorMapEncodeTiled (5-D patch map) failed with CUresult %d", (int)r);
    tmap_cache_put(key, *map);
    return LL_OK;
}

// optional fused-LayerNorm epilogue operands of the next gemm_launch on this thread
static thread_local const float2 *t_stats_in = nullptr;
static thread_local float2 *t_stats_out = nullptr;
static thread_local int t_stats_in_parts = 0;
static thread_local int t_stat_width = 1;

template <typename T, int BN, int kCtas, bool kDual = false>
static int launch_impl(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &td, const CUtensorMap &tr,
                       const float *bias, int has_residual, int64_t M, int N, int K, int act, cudaStream_t stream)
{
    using S = GemmShape<BN, kCtas>;
    static bool configured = false;
    auto kern = gemm_tcgen05_kernel<T, BN, kCtas, kDual>;
    if (!configured) {
        LL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmemBytes));
        configured = true;
    }
    int tiles = (int)((M + kBM * kCtas - 1) / (kBM * kCtas)) * (N / BN);
    if ((act & 0x10000) && !(act & 0x8000)) tiles *= t_stats_in_parts;  // split-K: (tile, split) work items
    const int workers = num_sms() / kCtas;
    const int grid = (tiles < workers ? tiles : workers) * kCtas;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = S::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    LL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tw, td, tr, bias, has_residual, (int)M, N, K, act, t_stats_in, t_stats_in_parts,
                                     t_stats_out, 1.f / (float)t_stat_width));
    return LL_OK;
}

// Tile choice by a small time model (microseconds): the persistent grid runs ceil(tiles / workers) rounds of one tile
// each, and a round costs a fixed part (pipeline fill, epilogue drain of the last tile: ~1.5 us) plus K / 64 K-blocks at a
// per-block time that depends on the tile shape.  The per-block times are measured on this kernel at the encoder's
// shapes (tools/gemm_bench.py, 32 ... 1512 images per pass, profiles/r02_gemm_small_batches.txt): a CTA of a pair stages
// half of the W tile, so 256 x 256 pair tiles do the most work per SM and per staged byte; 128 x 64 single-CTA tiles
// the least.  (Round 1 ranked candidates by rounds x tile width alone and never paired CTAs below one full round of
// pair tiles: at 128 images that chose 128 x 64 tiles for the two N = 768 GEMMs -- 86 us for c_proj where 256 x 192
// pair tiles take 45 -- and 7 rounds of 128 x 128 tiles for the in-projection where 4 rounds of 256 x 256 do.)
// The result does not depend on the choice: an output element's accumulation order over K is the same for every
// tile shape.
static bool g_splitk_enabled = true;  // ll_gemm_set_splitk (A/B knob)
static int g_force_bn = 0;    // test / tuning overrides (ll_gemm_set_tile_n, ll_gemm_set_cta_pair)
static int g_pair_mode = -1;  // -1 auto, 0 never, 1 always

static double kblock_us(int ctas, int bn)
{
    if (ctas == 2) return bn == 256 ? 0.44 : bn == 192 ? 0.39 : 0.36;
    return bn == 256 ? 0.51 : bn == 192 ? 0.40 : bn == 128 ? 0.31 : 0.30;
}

static void pick_tile(int64_t M, int N, int K, int &bn_out, int &ctas_out, bool mn_major_b = false)
{
    const int sms = num_sms();
    int best_bn = 0, best_ctas = 1;
    double best_cost = 0.0;
    const int cands[4] = {256, 192, 128, 64};
    const double num_kb = (double)((K + kBK - 1) / kBK);
    for (int ctas = 2; ctas >= 1; --ctas) {
        if (ctas == 2 && (g_pair_mode == 0 || (g_pair_mode < 0 && M <= kBM))) continue;  // a pair needs two row blocks
        if (ctas == 1 && g_pair_mode == 1) continue;
        const int64_t mt = (M + kBM * ctas - 1) / (kBM * ctas);
        const int workers = sms / ctas;
        for (int bn : cands) {
            if (N % bn != 0) continue;
            if (g_force_bn && bn != g_force_bn && N % g_force_bn == 0) continue;
            if (ctas == 2 && bn == 64) continue;  // 32-row W halves are not worth a pair
            if (mn_major_b && ctas == 2 && bn == 192) continue;  // a CTA of a pair stages whole 64-column atoms of an MN-major B
            const int64_t tiles = mt * (N / bn);
            const int64_t rounds = (tiles + workers - 1) / workers;
            const double cost = (double)rounds * (1.5 + num_kb * kblock_us(ctas, bn));
            if (best_bn == 0 || cost < best_cost - 1e-9) best_bn = bn, best_ctas = ctas, best_cost = cost;
        }
    }
    bn_out = best_bn;
    ctas_out = best_ctas;
}

template <typename T, int kCtas>
static int launch_bn(int bn, const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &td, const CUtensorMap &tr,
                     const float *bias, int has_residual, int64_t M, int N, int K, int act, cudaStream_t stream)
{
    if ((act & 0x80000) && !(act & 0x8000)) {  // dual output: its own instantiations
        if (has_residual) return fail(LL_ERR_INVALID, "gemm: dual output and a residual operand exclude each other");
        switch (bn) {
            case 256: return launch_impl<T, 256, kCtas, true>(ta, tw, td, tr, bias, 0, M, N, K, act, stream);
            case 192: return launch_impl<T, 192, kCtas, true>(ta, tw, td, tr, bias, 0, M, N, K, act, stream);
            case 128: return launch_impl<T, 128, kCtas, true>(ta, tw, td, tr, bias, 0, M, N, K, act, stream);
            case 64:
                if constexpr (kCtas == 1) return launch_impl<T, 64, 1, true>(ta, tw, td, tr, bias, 0, M, N, K, act, stream);
        }
        return fail(LL_ERR_INVALID, "gemm: unsupported tile N %d for %d CTA(s)", bn, kCtas);
    }
    switch (bn) {
        case 256: return launch_impl<T, 256, kCtas>(ta, tw, td, tr, bias, has_residual, M, N, K, act, stream);
        case 192: return launch_impl<T, 192, kCtas>(ta, tw, td, tr, bias, has_residual, M, N, K, act, stream);
        case 128: return launch_impl<T, 128, kCtas>(ta, tw, td, tr, bias, has_residual, M, N, K, act, stream);
        case 64:
            if constexpr (kCtas == 1) return launch_impl<T, 64, 1>(ta, tw, td, tr, bias, has_residual, M, N, K, act, stream);
    }
    return fail(LL_ERR_INVALID, "gemm: unsupported tile N %d for %d CTA(s)", bn, kCtas);
}

int gemm_launch_tmaps(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &td, const CUtensorMap &tr, int bn,
                      int ctas, const float *bias, int has_residual, int64_t M, int N, int K, int dtype, int act,
                      cudaStream_t stream)
{
    if (dtype == LL_F16)
        return ctas == 2 ? launch_bn<__half, 2>(bn, ta, tw, td, tr, bias, has_residual, M, N, K, act, stream)
                         : launch_bn<__half, 1>(bn, ta, tw, td, tr, bias, has_residual, M, N, K, act, stream);
    return ctas == 2 ? launch_bn<__nv_bfloat16, 2>(bn, ta, tw, td, tr, bias, has_residual, M, N, K, act, stream)
                     : launch_bn<__nv_bfloat16, 1>(bn, ta, tw, td, tr, bias, has_residual, M, N, K, act, stream);
}


int gemm_launch_ln(const void *A, const void *W, const float *bias, const void *residual, void *D, int64_t M, int N, int K,
                   int dtype, int act, const void *stats_in, int stats_in_parts, void *stats_out, int stat_width,
                   cudaStream_t stream)
{
    LL_REQUIRE(stats_in == nullptr || (stats_in_parts > 0 && stats_in_parts <= 16), "gemm_ln: stats_in needs 1 <= stats_in_parts <= 16");
    t_stats_in = static_cast<const float2 *>(stats_in);
    t_stats_in_parts = stats_in_parts;
    t_stats_out = static_cast<float2 *>(stats_out);
    t_stat_width = stat_width > 0 ? stat_width : 1;
    const int rc = gemm_launch(A, W, bias, residual, D, M, N, K, dtype, act, stream);
    t_stats_in = nullptr;
    t_stats_out = nullptr;
    return rc;
}

int gemm_launch(const void *A, const void *W, const float *bias, const void *residual, void *D, int64_t M, int N,
                int K, int dtype, int act, cudaStream_t stream)
{
    LL_REQUIRE(A && W && D, "gemm: null pointer");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "gemm: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(M > 0 && N > 0 && K > 0 && K % kBK == 0 && N % 64 == 0, "gemm: need K %% 64 == 0 and N %% 64 == 0 (M=%lld N=%d K=%d)", (long long)M, N, K);
    LL_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0 && (!residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0) &&
                   (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
               "gemm: D / residual / bias must be 16-byte aligned");
    LL_REQUIRE((act & 0xff) != 2 || (residual != nullptr && residual != D), "gemm: act 2 (QuickGELU backward) needs the pre-activation as `residual`, distinct from D");
    int bn = 0, ctas = 1;
    pick_tile(M, N, K, bn, ctas);
    CUtensorMap ta, tw, td, tr;
    int rc = make_tmap_2d(&ta, A, dtype, M, K, kBM);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&tw, W, dtype, N, K, bn / ctas);  // a CTA of a pair stages half of the W tile
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&td, D, dtype, M, N, 32);  // epilogue staging box: 32 rows x 64 columns
    if (rc != LL_OK) return rc;
    tr = td;
    if (residual != nullptr && residual != D) {
        rc = make_tmap_2d(&tr, residual, dtype, M, N, 32);
        if (rc != LL_OK) return rc;
    }
    return gemm_launch_tmaps(ta, tw, td, tr, bn, ctas, bias, residual != nullptr, M, N, K, dtype, act, stream);
}

// U = A W^T + bias and F = QuickGELU(U) from ONE pass (c_fc of the fine-tuning forward keeps both).
int gemm_dual_launch(const void *A, const void *W, const float *bias, void *U, void *F, int64_t M, int N, int K, int dtype, cudaStream_t stream)
{
    LL_REQUIRE(A && W && U && F && U != F, "gemm_dual: null pointer / U and F must differ");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "gemm_dual: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(M > 0 && N > 0 && K > 0 && K % kBK == 0 && N % 64 == 0, "gemm_dual: need K %% 64 == 0 and N %% 64 == 0 (M=%lld N=%d K=%d)", (long long)M, N, K);
    LL_REQUIRE(((reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(F)) & 15) == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
               "gemm_dual: U / F / bias must be 16-byte aligned");
    int bn = 0, ctas = 1;
    pick_tile(M, N, K, bn, ctas);
    CUtensorMap ta, tw, td, tr;
    int rc = make_tmap_2d(&ta, A, dtype, M, K, kBM);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&tw, W, dtype, N, K, bn / ctas);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&td, U, dtype, M, N, 32);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&tr, F, dtype, M, N, 32);
    if (rc != LL_OK) return rc;
    return gemm_launch_tmaps(ta, tw, td, tr, bn, ctas, bias, 0, M, N, K, dtype, 1 | 0x80000, stream);
}

// D[M, N] = act(A[M, K] @ B[K, N] + bias) (+ residual): B row-major [K, N] (MN-major operand).  The data gradient of a
// linear layer, dX = dY @ W with W as nn.Linear stores it -- no transposed weight copy.
int gemm_nn_launch(const void *A, const void *B, const float *bias, const void *residual, void *D, int64_t M, int N, int K, int dtype, int act,
                   cudaStream_t stream)
{
    LL_REQUIRE(A && B && D, "gemm_nn: null pointer");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "gemm_nn: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(M > 0 && N > 0 && K > 0 && K % kBK == 0 && N % 64 == 0, "gemm_nn: need K %% 64 == 0 and N %% 64 == 0 (M=%lld N=%d K=%d)", (long long)M, N, K);
    LL_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0 && (!residual || (reinterpret_cast<uintptr_t>(residual) & 15) == 0) &&
                   (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0),
               "gemm_nn: D / residual / bias must be 16-byte aligned");
    LL_REQUIRE((act & 0xff) != 2 || (residual != nullptr && residual != D), "gemm_nn: act 2 (QuickGELU backward) needs the pre-activation as `residual`, distinct from D");
    int bn = 0, ctas = 1;
    pick_tile(M, N, K, bn, ctas, true);
    CUtensorMap ta, tw, td, tr;
    int rc = make_tmap_2d(&ta, A, dtype, M, K, kBM);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&tw, B, dtype, K, N, 64);  // 64-row (K) x 64-column (N) boxes
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&td, D, dtype, M, N, 32);
    if (rc != LL_OK) return rc;
    tr = td;
    if (residual != nullptr && residual != D) {
        rc = make_tmap_2d(&tr, residual, dtype, M, N, 32);
        if (rc != LL_OK) return rc;
    }
    return gemm_launch_tmaps(ta, tw, td, tr, bn, ctas, bias, residual != nullptr, M, N, K, dtype, (act & 0xff) | 0x40000, stream);
}

int gemm_patch_launch(const void *x, const void *W, void *D, int64_t n_images, int N, int res, int patch, int dtype, cudaStream_t stream)
{
    LL_REQUIRE(x && W && D, "gemm_patch: null pointer");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "gemm_patch: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(patch == 32 && res % patch == 0 && res / patch <= kPatchGridPitch && n_images > 0 && N % 64 == 0,
               "gemm_patch: needs patch 32, a patch grid of at most 8 x 8 and N %% 64 == 0 (res=%d patch=%d N=%d)", res, patch, N);
    LL_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0, "gemm_patch: D must be 16-byte aligned");
    const int64_t M = n_images * kPatchRowsPerImage;
    const int K = 3 * patch * patch;
    int bn = 0, ctas = 1;
    pick_tile(M, N, K, bn, ctas);
    CUtensorMap ta, tw, td;
    int rc = make_tmap_patch5d(&ta, x, dtype, n_images, res, patch);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&tw, W, dtype, N, K, bn / ctas);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&td, D, dtype, M, N, 32);
    if (rc != LL_OK) return rc;
    const int kb_per_channel = patch * patch / kBK;
    return gemm_launch_tmaps(ta, tw, td, td, bn, ctas, nullptr, 0, M, N, K, dtype, 0x8000 | (kb_per_channel << 16), stream);
}

// D[i] = T(sum_s ws[s][i]): adds the float32 partials of a split-K launch, one rounding to the 16-bit result
template <typename T>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float *__restrict__ ws, int splits, int64_t n8, int64_t plane, T *__restrict__ D)
{
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < splits; ++s) {
            const float4 a = *reinterpret_cast<const float4 *>(ws + s * plane + i * 8), b = *reinterpret_cast<const float4 *>(ws + s * plane + i * 8 + 4);
            acc[0] += a.x, acc[1] += a.y, acc[2] += a.z, acc[3] += a.w, acc[4] += b.x, acc[5] += b.y, acc[6] += b.z, acc[7] += b.w;
        }
        *reinterpret_cast<uint4 *>(D + i * 8) = Pack8<T>::pack(acc);
    }
}

// Shape of a split-K weight-gradient launch: CTA pairs or single CTAs, tile width, number of K splits.  Chosen by a small
// time model (microseconds) whose constants are the measured per-K-block times of the TN kernel at a 25 600-deep
// reduction: a 256 x 256 pair tile and a 128 x 128 single-CTA tile both take ~0.47 us per 64-deep K block (the single
// CTA is bound by its operand feed: 64 FLOP per staged byte against 128 for the pair), so pairs do twice the work per SM
// whenever the splits can fill 74 of them; every extra split costs one float32 partial of the output, written and read
// once (L2-resident at these sizes: <= 38 MB).
struct SplitPlan {
    int ctas, bn, splits;
    double us;
};

static SplitPlan plan_splitk(int Mo, int No, int64_t num_kb, int64_t workspace_bytes)
{
    const int sms = num_sms();
    SplitPlan best = {1, 0, 1, 1e30};
    for (int ctas = 2; ctas >= 1; --ctas) {
        if (ctas == 2 && (g_pair_mode == 0 || Mo % 256 != 0)) continue;
        if (ctas == 1 && g_pair_mode == 1) continue;
        const int workers = sms / ctas;
        const int bns[3] = {256, 128, 64};
        for (int bn : bns) {
            if (No % bn != 0 || (ctas == 2 && bn == 64)) continue;
            if (g_force_bn && bn != g_force_bn && No % g_force_bn == 0) continue;
            // per-K-block time of one work item
            const double t_kb = ctas == 2 ? (bn == 256 ? 0.47 : 0.30) : (bn == 256 ? 0.70 : bn == 128 ? 0.47 : 0.32);
            const int64_t tiles = (int64_t)(Mo / (kBM * ctas)) * (No / bn);
            for (int s = 1; s <= 16 && s <= num_kb; ++s) {
                if (s > 1 && (int64_t)s * Mo * No * (int64_t)sizeof(float) > workspace_bytes) break;
                const int64_t rounds = (tiles * s + workers - 1) / workers;
                double us = (double)rounds * (double)((num_kb + s - 1) / s) * t_kb + 4.0 * (double)rounds;
                if (s > 1) us += 3.0 + (double)s * Mo * No * 8.0 / 5.0e6;  // reduce launch + partials at ~5 TB/s
                if (us < best.us - 1e-9) best = {ctas, bn, s, us};
            }
        }
    }
    return best;
}

int64_t gemm_tn_workspace_bytes(int Mo, int No) { return 8LL * Mo * No * (int64_t)sizeof(float); }

// D[Mo, No] = A[Kt, Mo]^T @ B[Kt, No]: the weight-gradient contraction dW = dY^T X of a linear layer, with the
// reduction (token) dimension as the ROWS of both operands, i.e. exactly as the activations and their gradients
// sit in memory.  Same kernel, act bit 0x2000: MN-major shared-memory descriptors, 64 x 64 TMA boxes.
int gemm_tn_launch(const void *A, const void *B, void *D, int64_t Kt, int Mo, int No, int dtype, cudaStream_t stream, void *workspace,
                   int64_t workspace_bytes)
{
    LL_REQUIRE(A && B && D, "gemm_tn: null pointer");
    LL_REQUIRE(dtype == LL_F16 || dtype == LL_BF16, "gemm_tn: dtype must be LL_F16 or LL_BF16");
    LL_REQUIRE(Kt > 0 && Mo > 0 && No > 0 && Mo % 128 == 0 && No % 64 == 0, "gemm_tn: need Mo %% 128 == 0 and No %% 64 == 0 (Kt=%lld Mo=%d No=%d)", (long long)Kt, Mo, No);
    LL_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0, "gemm_tn: D must be 16-byte aligned");
    int bn = 0, ctas = 1;
    pick_tile(Mo, No, (int)std::min<int64_t>(Kt, 1 << 30), bn, ctas);
    if (ctas == 2 && (bn == 192 || Mo % 256 != 0)) {  // a pair stages BN / 2 columns of B per CTA: whole 64-column atoms only
        const int saved_mode = g_pair_mode;
        g_pair_mode = 0;  // the best single-CTA tile instead
        pick_tile(Mo, No, (int)std::min<int64_t>(Kt, 1 << 30), bn, ctas);
        g_pair_mode = saved_mode;
        ctas = 1;
        if (g_pair_mode == 1 && No % 256 == 0 && Mo % 256 == 0) ctas = 2, bn = 256;
    }
    CUtensorMap ta, tw, td;
    int rc = make_tmap_2d(&ta, A, dtype, Kt, Mo, 64);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&tw, B, dtype, Kt, No, 64);
    if (rc != LL_OK) return rc;
    rc = make_tmap_2d(&td, D, dtype, Mo, No, 32);
    if (rc != LL_OK) return rc;
    const int64_t Kpad = (Kt + kBK - 1) / kBK * kBK;
    LL_REQUIRE(Kpad < (1ll << 31), "gemm_tn: Kt too large");
    // split-K when the output is a fraction of a round of tiles and the caller lent a workspace: a [768..3072]^2 weight
    // gradient is 9-36 pair tiles of 256 x 256 for a 25 600-deep reduction
    if (workspace != nullptr && g_splitk_enabled && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && ((int64_t)Mo * No) % 8 == 0) {
        const SplitPlan plan = plan_splitk(Mo, No, Kpad / kBK, workspace_bytes);
        if (plan.bn != 0 && plan.splits > 1) {
            t_stats_in = nullptr;
            t_stats_in_parts = plan.splits;
            t_stats_out = static_cast<float2 *>(workspace);
            rc = gemm_launch_tmaps(ta, tw, td, td, plan.bn, plan.ctas, nullptr, 0, Mo, No, (int)Kpad, dtype, 0x2000 | 0x10000, stream);
            t_stats_in_parts = 0;
            t_stats_out = nullptr;
            if (rc != LL_OK) return rc;
            const int64_t n8 = (int64_t)Mo * No / 8;
            const unsigned grid = (unsigned)std::min<int64_t>((n8 + 255) / 256, 8LL * num_sms());
            if (dtype == LL_F16)
                LL_CUDA_CHECK(launch_pdl(splitk_reduce_kernel<__half>, dim3(grid), dim3(256), 0, stream, static_cast<const float *>(workspace), plan.splits, n8, (int64_t)Mo * No, static_cast<__half *>(D)));
            else
                LL_CUDA_CHECK(launch_pdl(splitk_reduce_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, static_cast<const float *>(workspace), plan.splits, n8, (int64_t)Mo * No, static_cast<__nv_bfloat16 *>(D)));
            return LL_OK;
        }
    }
    return gemm_launch_tmaps(ta, tw, td, td, bn, ctas, nullptr, 0, Mo, No, (int)Kpad, dtype, 0x2000, stream);
}

}  // namespace ll

extern "C" int ll_gemm_tn(const void *A, const void *B, void *D, int64_t Kt, int Mo, int No, int dtype, ll_stream_t stream)
{
    return ll::gemm_tn_launch(A, B, D, Kt, Mo, No, dtype, ll::as_stream(stream), nullptr, 0);
}

extern "C" int ll_gemm_dual(const void *A, const void *W, const float *bias, void *U, void *F, int64_t M, int N, int K, int dtype,
                            ll_stream_t stream)
{
    return ll::gemm_dual_launch(A, W, bias, U, F, M, N, K, dtype, ll::as_stream(stream));
}

extern "C" int ll_gemm_nn(const void *A, const void *B, const float *bias, const void *residual, void *D, int64_t M, int N, int K, int dtype,
                          int act, ll_stream_t stream)
{
    return ll::gemm_nn_launch(A, B, bias, residual, D, M, N, K, dtype, act, ll::as_stream(stream));
}

extern "C" int64_t ll_gemm_tn_workspace_bytes(int Mo, int No) { return ll::gemm_tn_workspace_bytes(Mo, No); }

extern "C" int ll_gemm_tn_splitk(const void *A, const void *B, void *D, int64_t Kt, int Mo, int No, int dtype, void *workspace,
                                 int64_t workspace_bytes, ll_stream_t stream)
{
    return ll::gemm_tn_launch(A, B, D, Kt, Mo, No, dtype, ll::as_stream(stream), workspace, workspace_bytes);
}

extern "C" int ll_gemm_set_splitk(int on)
{
    ll::g_splitk_enabled = on != 0;
    return LL_OK;
}

extern "C" int ll_gemm_patch_embed(const void *x, const void *W, void *D, int64_t n_images, int N, int res, int patch, int dtype,
                                   ll_stream_t stream)
{
    return ll::gemm_patch_launch(x, W, D, n_images, N, res, patch, dtype, ll::as_stream(stream));
}

extern "C" int ll_gemm_set_tile_n(int bn)
{
    if (bn != 0 && bn != 64 && bn != 128 && bn != 192 && bn != 256) return ll::fail(LL_ERR_INVALID, "gemm_set_tile_n: %d not in {0,64,128,192,256}", bn);
    ll::g_force_bn = bn;
    return LL_OK;
}

extern "C" int ll_gemm_set_cta_pair(int mode)
{
    if (mode < -1 || mode > 1) return ll::fail(LL_ERR_INVALID, "gemm_set_cta_pair: mode must be -1 (auto), 0 (off) or 1 (on)");
    ll::g_pair_mode = mode;
    return LL_OK;
}

extern "C" int ll_gemm_ln(const void *A, const void *W, const float *bias, const void *residual, void *D, int64_t M, int N,
                          int K, int dtype, int act, const void *row_stats_in, int stats_in_parts, void *row_stats_out,
                          int stat_width, ll_stream_t stream)
{
    return ll::gemm_launch_ln(A, W, bias, residual, D, M, N, K, dtype, act, row_stats_in, stats_in_parts, row_stats_out,
                              stat_width, ll::as_stream(stream));
}

extern "C" int ll_gemm(const void *A, const void *W, const float *bias, const void *residual, void *D, int64_t M,
                       int N, int K, int dtype, int act, ll_stream_t stream)
{
    return ll::gemm_launch(A, W, bias, residual, D, M, N, K, dtype, act, ll::as_stream(stream));
}
