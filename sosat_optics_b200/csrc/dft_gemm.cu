// Stage 2 of the hot path on sm_100a: the far-field transform of opt_analyze.a2b
// (opt_analyze.py:220-230),  B = fftshift(fft2(fftshift(b)))  ==  W b W^T  with
//     W[k, n] = exp(-2 pi i ((k - floor(N/2)) (n - ceil(N/2)) mod N) / N),
// executed as a dense separable complex contraction: two real GEMM passes on tcgen05 tensor
// cores (UTCHMMA), operands staged by TMA into 128B-swizzled shared memory, fp32 accumulators
// in TMEM.  A single bf16 pass is not accurate enough for the 1e-4-of-peak tolerance
// (SURVEY section 7), so every operand is split x = hi + lo (two bf16) and each k-block issues
// the three products hi*hi + hi*lo + lo*hi into the same accumulator ("bf16x3").
//
//   prep:    X[m, (n,c')] = b[n, m]                 (transpose + bf16 hi/lo split, zero padded)
//   pass 1:  D1[(l,c), m] = sum_n  W[(l,c), (n,c')] * X[m, (n,c')]     = (W b)[l, m]
//   pass 2:  B[l, (k,c)]  = sum_m  D1[l, (c',m)]    * W[(k,c), (c',m)] = (W b W^T)[l, k]
//
// (.,c) = real/imag component.  The complex structure is folded into the twiddle operand:
// row (l,re) = [Wr, -Wi], row (l,im) = [Wi, Wr] along K, so the data operand is just the
// complex array viewed as reals.  Both passes are "TN" GEMMs with K-major operands; pass 1
// writes its result already split into bf16 hi/lo in exactly the layout pass 2 reads, and pass 2
// (twiddle on the N side) leaves the accumulator rows as rows of B with interleaved re/im
// columns, so its epilogue stores the complex64 result directly -- no finishing pass.
//
// The mainloop of a 128 x 128 tile needs 64 KB of operands per 12 UMMAs (768 tensor-pipe
// cycles): 85 B/clk/SM, more than L2 can feed 128-148 SMs at once (the first cut was
// L2->SM bandwidth bound: 25 us per 1024^2 pass against 12.5 us of UMMA time).  CTAs are
// therefore launched as CX x CY clusters: the CX CTAs of a cluster row share the A tile and
// the CY CTAs of a column share the B tile; each CTA TMA-loads a 1/CX (1/CY) slice and
// multicasts it, which divides the L2 traffic per CTA by 2 for 2 x 2 clusters.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace sosat {

// CTA tile 128 x 128; a k-block is one swizzle row: BK = 64 (128 B swizzle, 3 stages of 64 KB) or
// BK = 32 (64 B swizzle, 6 stages of 32 KB: the same 192 KB in flight in finer units, so that a
// freed stage leaves 5/6 instead of 2/3 of the ring to cover the refill latency).
#ifndef SOSAT_GEMM_BK
#define SOSAT_GEMM_BK 64
#endif
constexpr int BM = 128, BN = 128, BK = SOSAT_GEMM_BK;
static_assert(BK == 64 || BK == 32, "BK must be one 128 B or 64 B swizzle row of bf16");
constexpr int UMMA_K = 16;
constexpr int kStages = BK == 64 ? 3 : 6;
constexpr int kRowBytes = BK * 2;                    // one K-major row of a tile
constexpr int kTileBytes = BM * BK * 2;              // 16 KiB per operand tile
constexpr int kStageBytes = 4 * kTileBytes;          // A_hi, A_lo, B_hi, B_lo
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kTmemCols = 128;                       // 128 lanes x 128 fp32 columns

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                               int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// programmatic dependent launch: wait for the prerequisite grid (and its memory), and let the
// dependent grid start its prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(bar) : "memory");
}
// commit that arrives on the barrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, swizzled operand tile: rows of kRowBytes (128 B or 64 B = the swizzle span), 8-row
// swizzle atoms 8 * kRowBytes apart (SBO; LBO unused for swizzled K-major layouts and set to 1),
// descriptor version 1, layout type 2 = SWIZZLE_128B / 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) |
           ((uint64_t)((8 * kRowBytes) >> 4) << 32) | (1ull << 46) |
           ((uint64_t)(BK == 64 ? 2 : 4) << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

enum { EPI_C64 = 0, EPI_SPLIT_BF16 = 1, EPI_RED_C64 = 2, EPI_RED_F32 = 3 };

// TMEM accumulator tile (128 lanes x 128 fp32 columns) -> registers -> global.
//   EPI_SPLIT_BF16: bf16 hi/lo planes d_hi/d_lo [M][ldd]
//   EPI_C64:        times scale -> d_c64 [n_valid][2 n_valid] floats = complex64 [n][n]
//   EPI_RED_C64:    the same with red.global.add (split-K partial sums into a zeroed output)
//   EPI_RED_F32:    red.global.add.v4.f32 into the zeroed dense fp32 matrix d_c64 [M][ldd]
// `partial` (optional): the other K slice's fp32 sums of this CTA's tile, stored by
// store_partial_tile as float4 [NCOLS/4][128 rows] (row fastest: coalesced both ways); added in.
template <int EPI, int NCOLS = BN>
__device__ __forceinline__ void gemm_epilogue(uint32_t tmem_base, int m0, int n0, int warp, int lane,
                                              float *__restrict__ d_c64,
                                              __nv_bfloat16 *__restrict__ d_hi,
                                              __nv_bfloat16 *__restrict__ d_lo, int ldd, int n_valid,
                                              float scale, const float4 *partial = nullptr) {
    const int row = m0 + warp * 32 + lane;  // TMEM lane == accumulator row
#pragma unroll 1
    for (int c = 0; c < NCOLS / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        if (partial) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 o = partial[(size_t)(c * 8 + q) * BM + warp * 32 + lane];
                v[4 * q] = __float_as_uint(__uint_as_float(v[4 * q]) + o.x);
                v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + o.y);
                v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + o.z);
                v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + o.w);
            }
        }
        if (EPI == EPI_RED_F32) {
            float4 *dst = (float4 *)(d_c64 + (size_t)row * ldd + n0 + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                atomicAdd(dst + q, make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                               __uint_as_float(v[4 * q + 2]),
                                               __uint_as_float(v[4 * q + 3])));
        } else if (EPI == EPI_C64 || EPI == EPI_RED_C64) {
            // row = l, columns = (k, re|im) interleaved: the complex64 row of B, n_valid wide
            const int j0 = n0 + c * 32;
            if (row < n_valid && j0 < 2 * n_valid) {
                float *dst = d_c64 + (size_t)row * (2 * (size_t)n_valid) + j0;
                if ((n_valid & 1) == 0 && j0 + 32 <= 2 * n_valid) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 o = make_float4(scale * __uint_as_float(v[4 * q]),
                                                     scale * __uint_as_float(v[4 * q + 1]),
                                                     scale * __uint_as_float(v[4 * q + 2]),
                                                     scale * __uint_as_float(v[4 * q + 3]));
                        if (EPI == EPI_RED_C64) atomicAdd((float4 *)dst + q, o);
                        else ((float4 *)dst)[q] = o;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (j0 + 2 * q + 1 < 2 * n_valid) {
                            const float2 o = make_float2(scale * __uint_as_float(v[2 * q]),
                                                         scale * __uint_as_float(v[2 * q + 1]));
                            if (EPI == EPI_RED_C64) atomicAdd((float2 *)dst + q, o);
                            else ((float2 *)dst)[q] = o;
                        }
                }
            }
        } else {
            const size_t off = (size_t)row * ldd + n0 + c * 32;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float x0 = __uint_as_float(v[2 * q]), x1 = __uint_as_float(v[2 * q + 1]);
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                hi[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lo[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            uint4 *dh = (uint4 *)(d_hi + off), *dl = (uint4 *)(d_lo + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                dh[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                dl[q] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
        }
    }
}

template <int NCOLS>
__device__ __forceinline__ void store_partial_tile(uint32_t tmem_base, int warp, int lane,
                                                   float4 *__restrict__ partial) {
#pragma unroll 1
    for (int c = 0; c < NCOLS / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            partial[(size_t)(c * 8 + q) * BM + warp * 32 + lane] =
                make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                            __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
}

// D[M][N] (+)= A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T ; one 128 x 128 tile per CTA, CTAs in
// CX x CY clusters (cluster x = blockIdx.x = N tiles, y = M tiles).
// warp 0 lane 0: TMA producer.  warp 1 lane 0: MMA issuer.  all 4 warps: epilogue.
//   EPI_SPLIT_BF16: D -> bf16 hi/lo planes d_hi/d_lo [M][ldd]                       (pass 1)
//   EPI_C64:        D * scale -> d_c64 [n_valid][2 n_valid] floats = complex64 [n][n] (pass 2)
template <int EPI, int CX, int CY>
__global__ void __launch_bounds__(128, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                   const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi,
                   const __grid_constant__ CUtensorMap map_b_lo, int num_kb, float *__restrict__ d_c64,
                   __nv_bfloat16 *__restrict__ d_hi, __nv_bfloat16 *__restrict__ d_lo, int ldd,
                   int n_valid, float scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStages * kStageBytes);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages),
                   accum_bar = smem_u32(bars + 2 * kStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    constexpr bool kCluster = CX * CY > 1;
    constexpr int kSharers = CX + CY - 1;  // CTAs whose smem one CTA's loads (and commits) reach
    const uint32_t rank = kCluster ? cluster_ctarank() : 0u;
    const int cx = (int)rank % CX, cy = (int)rank / CX;
    uint16_t mask_a = 0, mask_b = 0;  // cluster row-mates share A, column-mates share B
#pragma unroll
    for (int x = 0; x < CX; ++x) mask_a |= (uint16_t)(1u << (x + cy * CX));
#pragma unroll
    for (int y = 0; y < CY; ++y) mask_b |= (uint16_t)(1u << (cx + y * CX));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, kSharers);  // every CTA that reads this stage's data
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();  // lane 0 rejoins before the warp-aligned allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (kCluster) cluster_sync_all();  // peers' barriers are initialised before anything is multicast
    pdl_wait();                        // operands come from the previous kernel in the stream
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        // ---- TMA producer: this CTA's slice of A (rows cx*128/CX ..) and of B, multicast ----
        constexpr int kRowsA = BM / CX, kRowsB = BN / CY;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t full = full0 + 8 * s;
            const uint32_t base = smem_u32(smem + s * kStageBytes);
            mbar_expect_tx(full, kStageBytes);  // own slices + the peers' multicasts
            const uint32_t da = base + (uint32_t)(cx * kRowsA * kRowBytes), db = base + (uint32_t)(cy * kRowsB * kRowBytes);
            if (CX > 1) {
                tma_load_2d_mc(da + 0 * kTileBytes, &map_a_hi, full, kb * BK, m0 + cx * kRowsA, mask_a);
                tma_load_2d_mc(da + 1 * kTileBytes, &map_a_lo, full, kb * BK, m0 + cx * kRowsA, mask_a);
            } else {
                tma_load_2d(da + 0 * kTileBytes, &map_a_hi, full, kb * BK, m0);
                tma_load_2d(da + 1 * kTileBytes, &map_a_lo, full, kb * BK, m0);
            }
            if (CY > 1) {
                tma_load_2d_mc(db + 2 * kTileBytes, &map_b_hi, full, kb * BK, n0 + cy * kRowsB, mask_b);
                tma_load_2d_mc(db + 3 * kTileBytes, &map_b_lo, full, kb * BK, n0 + cy * kRowsB, mask_b);
            } else {
                tma_load_2d(db + 2 * kTileBytes, &map_b_hi, full, kb * BK, n0);
                tma_load_2d(db + 3 * kTileBytes, &map_b_lo, full, kb * BK, n0);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer (single thread) ----
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(smem + s * kStageBytes);
            const uint64_t a_hi = umma_desc_sw128(base + 0 * kTileBytes);
            const uint64_t a_lo = umma_desc_sw128(base + 1 * kTileBytes);
            const uint64_t b_hi = umma_desc_sw128(base + 2 * kTileBytes);
            const uint64_t b_lo = umma_desc_sw128(base + 3 * kTileBytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);  // +32 B per K=16 step
                umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
                umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
            }
            // stage reusable once these MMAs retire: tell every CTA that loads into this one
            if (kCluster) umma_commit_mc(empty0 + 8 * s, (uint16_t)(mask_a | mask_b));
            else umma_commit(empty0 + 8 * s);
        }
        umma_commit(accum_bar);
    }
    __syncwarp();

    // ---- epilogue: TMEM -> registers -> global ----
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    gemm_epilogue<EPI>(tmem_base, m0, n0, warp, lane, d_c64, d_hi, d_lo, ldd, n_valid, scale);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(kTmemCols) : "memory");
    // no CTA may exit while a peer can still multicast into its shared memory or barriers
    if (kCluster) cluster_sync_all();
}

// ---- CTA-pair variant: tcgen05.mma.cta_group::2 on a 256 x 128 pair tile ---------------------
// Two CTAs of a (2, 1, 1) cluster (adjacent M tiles = blockIdx.x, same N tile = blockIdx.y; the
// pair must be the cluster's x dimension) execute ONE UMMA of M = 256:
// each CTA holds its own 128 rows of A and HALF of the B tile (64 of the 128 rows); the tensor
// cores of the two SMs exchange the B halves, so a CTA's shared memory sees 3 x (4 KB A + 2 KB B)
// of operand reads per 192 tensor cycles (96 B/clk instead of 128) and 48 KB instead of 64 KB of
// TMA fill per k-block (62 B/clk instead of 85) -- and the ring is 4 stages deep in the same
// 192 KB.  Roles: both CTAs run a TMA producer (their A rows, their B half; the bytes are
// accounted on the LEADER's full barrier through the .cta_group::2 TMA form), only the leader
// (cluster rank 0) issues the UMMAs, and its tcgen05.commit is multicast to the stage-free and
// accumulator-ready barriers of both CTAs.  Each CTA drains its own 128 TMEM lanes.
// BN2 = N of the pair tile: 128 (4 stages of 48 KB) or 256 (3 stages of 64 KB, 256 TMEM columns).
template <int BN2> struct PairCfg {
    static constexpr int kHalfBytes = (BN2 / 2) * BK * 2;              // half of a B tile
    static constexpr int kStageBytes = 2 * kTileBytes + 2 * kHalfBytes;  // per CTA and stage
    static constexpr int kStages = BN2 == 128 ? 4 : 3;
    static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
    static constexpr int kTmem = BN2;
};

__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar)
                 : "memory");
}
// TMA load whose completion bytes are signalled on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map,
                                                 uint32_t cluster_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar), "h"(cta_mask) : "memory");
}

// Batched second pass (sosat_dft2_gemm_batch): the M tiles of all batch items are laid along
// grid x, every item owns `mtiles` of them; the A operand of item i is the slice
// D1[l][(c', i np + m)] of the pass-1 output of the whole batch, so k-block kb reads columns
// (kb / kpc) seg + i kpc 64 + (kb % kpc) 64 of the [np][2 seg] view; the output of item i goes to
// d_c64 + i out_stride.  All zero = the plain single-transform GEMM.
struct PairBatch {
    int kpc;        // k-blocks per c' segment (np / 64)
    int seg;        // columns per c' segment of the D1 view (batch * np)
    int mtiles;     // M tiles per batch item (np / 128)
    long long out_stride;  // floats between the outputs of consecutive items (2 n^2)
};

#ifdef SOSAT_EXPERIMENTS  // one-tile-per-CTA form of the pair kernel (A/B runs: SOSAT_GEMM_PERSIST=0, SOSAT_GEMM_PAIR)
template <int EPI, int BN2>
__global__ void __launch_bounds__(128, 1)
gemm2_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                    const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi,
                    const __grid_constant__ CUtensorMap map_b_lo, int num_kb, float *__restrict__ d_c64,
                    __nv_bfloat16 *__restrict__ d_hi, __nv_bfloat16 *__restrict__ d_lo, int ldd,
                    int n_valid, float scale, PairBatch pb) {
    using Cfg = PairCfg<BN2>;
    constexpr int kStages2 = Cfg::kStages, kStageBytes2 = Cfg::kStageBytes,
                  kHalfTileBytes = Cfg::kHalfBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStages2 * kStageBytes2);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages2 + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages2),
                   accum_bar = smem_u32(bars + 2 * kStages2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0: leader (even M tile), 1: peer
    // Rasterisation: the pair must lie along grid x, but consecutive pairs should sweep the N
    // tiles of one pair of A row-tiles (as the single-CTA kernel does) -- with M fastest a wave
    // touches the whole twiddle matrix and DRAM traffic doubles (measured 280 vs 126 MB at 2048^2).
    const unsigned pair_id = (blockIdx.x >> 1) + (gridDim.x >> 1) * blockIdx.y;
    const int n0 = (int)(pair_id % gridDim.y) * BN2;
    const int m_tile = (int)(2 * (pair_id / gridDim.y) + rank);  // this CTA's own 128 rows of D
    const int item = pb.mtiles ? m_tile / pb.mtiles : 0;
    const int m0 = (m_tile - item * pb.mtiles) * BM;             // rows within the batch item
    const int a_col0 = item * pb.kpc * BK;
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(full0 + 8 * s, 2);   // leader: its own arrive.expect_tx + the peer's arrive
            mbar_init(empty0 + 8 * s, 1);  // the leader's multicast commit
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(Cfg::kTmem) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();  // both CTAs' barriers and TMEM are set up
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        // ---- TMA producer (both CTAs): own A rows, own half of B; bytes land on the leader ----
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages2;
            const uint32_t ph = (uint32_t)(kb / kStages2) & 1u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t lead_full = map_to_cta(full0 + 8 * s, 0);
            const uint32_t base = smem_u32(smem + s * kStageBytes2);
            if (leader) mbar_expect_tx(full0 + 8 * s, 2 * kStageBytes2);
            else mbar_arrive_cluster(lead_full);
            const int a_col = pb.kpc ? (kb / pb.kpc) * pb.seg + a_col0 + (kb % pb.kpc) * BK : kb * BK;
            tma_load_2d_pair(base, &map_a_hi, lead_full, a_col, m0);
            tma_load_2d_pair(base + kTileBytes, &map_a_lo, lead_full, a_col, m0);
            tma_load_2d_pair(base + 2 * kTileBytes, &map_b_hi, lead_full, kb * BK,
                             n0 + (int)rank * (BN2 / 2));
            tma_load_2d_pair(base + 2 * kTileBytes + kHalfTileBytes, &map_b_lo, lead_full, kb * BK,
                             n0 + (int)rank * (BN2 / 2));
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ---- MMA issuer: one thread of the leader CTA drives both SMs' tensor cores ----
        constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN2);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages2;
            const uint32_t ph = (uint32_t)(kb / kStages2) & 1u;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(smem + s * kStageBytes2);
            const uint64_t a_hi = umma_desc_sw128(base);
            const uint64_t a_lo = umma_desc_sw128(base + kTileBytes);
            const uint64_t b_hi = umma_desc_sw128(base + 2 * kTileBytes);
            const uint64_t b_lo = umma_desc_sw128(base + 2 * kTileBytes + kHalfTileBytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                umma2_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
                umma2_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                umma2_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
            }
            umma2_commit_mc(empty0 + 8 * s, (uint16_t)0x3);  // frees the stage in both CTAs
        }
        umma2_commit_mc(accum_bar, (uint16_t)0x3);
    }
    __syncwarp();

    // ---- epilogue (both CTAs): own 128 TMEM lanes -> global ----
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    gemm_epilogue<EPI, BN2>(tmem_base, m0, n0, warp, lane,
                            d_c64 ? d_c64 + (long long)item * pb.out_stride : nullptr, d_hi, d_lo, ldd,
                            n_valid, scale);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(Cfg::kTmem) : "memory");
}

#endif  // SOSAT_EXPERIMENTS
// ---- persistent CTA-pair kernel: one pair per TPC, two TMEM accumulator stages -----------------
// The 256 x 256 pair-tile kernel above as a persistent loop: a pair walks the pair tiles
// t = pair, pair + P, ... and the accumulator alternates between two 256-column TMEM stages, so
// that the epilogue of tile i (4 dedicated warps per CTA: TMEM -> registers -> global) runs under
// the mainloop of tile i + 1, and barrier set-up / TMEM allocation / cluster sync happen once per
// SM instead of once per tile.  Warps: 0 TMA producer, 1 MMA issuer (leader CTA only), 2-5
// epilogue (warp w drains TMEM lanes 32 (w & 3) ..).  Barriers per CTA: smem ring full/empty as
// above; tmem_full[a] (the leader's multicast commit, count 1); tmem_empty[a] on the LEADER
// (count 8: one lane of each epilogue warp of both CTAs, the peer's through mapa).
constexpr int kPersistThreads = 192;

template <int EPI>
__global__ void __launch_bounds__(kPersistThreads, 1)
gemm2p_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                     const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi,
                     const __grid_constant__ CUtensorMap map_b_lo, int num_kb, int m_pairs,
                     int n_tiles, int ksplit, float *__restrict__ d_c64,
                     __nv_bfloat16 *__restrict__ d_hi, __nv_bfloat16 *__restrict__ d_lo, int ldd,
                     int n_valid, float scale, PairBatch pb) {
    constexpr int BN2 = 256;
    using Cfg = PairCfg<BN2>;
    constexpr int kStages2 = Cfg::kStages, kStageBytes2 = Cfg::kStageBytes,
                  kHalfTileBytes = Cfg::kHalfBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStages2 * kStageBytes2);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages2 + 4);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages2),
                   tfull0 = smem_u32(bars + 2 * kStages2), tempty0 = smem_u32(bars + 2 * kStages2 + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs_resident = gridDim.x >> 1;
    // work item = (pair tile, K slice): with ksplit > 1 the slices of a tile run on different
    // pairs and meet in the output through red.global.add (EPI_RED_*), so that sizes with fewer
    // pair tiles than TPCs (1024^2: 32) still fill the machine
    const int total_tiles = m_pairs * n_tiles * ksplit;
    const int kb_per = num_kb / ksplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(full0 + 8 * s, 2);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();
    pdl_wait();
    pdl_launch_dependents();

    // tile t of this pair -> (m pair, n tile), N fastest (see the rasterisation note above)
    auto tile_coords = [&](int w, int &m0, int &n0, int &item, int &a_col0) {
        const int t = w / ksplit;
        n0 = (t % n_tiles) * BN2;
        const int m_tile = 2 * (t / n_tiles) + (int)rank;
        item = pb.mtiles ? m_tile / pb.mtiles : 0;
        m0 = (m_tile - item * pb.mtiles) * BM;
        a_col0 = item * pb.kpc * BK;
    };

    if (warp == 0 && lane == 0) {
        // ---- TMA producer ----
        uint32_t g = 0;  // k-blocks issued so far: ring position and phase run across tiles
        for (int t = pair; t < total_tiles; t += n_pairs_resident) {
            int m0, n0, item, a_col0;
            tile_coords(t, m0, n0, item, a_col0);
            const int kb0 = (t % ksplit) * kb_per;
            for (int kb = kb0; kb < kb0 + kb_per; ++kb, ++g) {
                const uint32_t s = g % kStages2, ph = (g / kStages2) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                const uint32_t lead_full = map_to_cta(full0 + 8 * s, 0);
                const uint32_t base = smem_u32(smem + s * kStageBytes2);
                if (leader) mbar_expect_tx(full0 + 8 * s, 2 * kStageBytes2);
                else mbar_arrive_cluster(lead_full);
                const int a_col = pb.kpc ? (kb / pb.kpc) * pb.seg + a_col0 + (kb % pb.kpc) * BK : kb * BK;
                tma_load_2d_pair(base, &map_a_hi, lead_full, a_col, m0);
                tma_load_2d_pair(base + kTileBytes, &map_a_lo, lead_full, a_col, m0);
                tma_load_2d_pair(base + 2 * kTileBytes, &map_b_hi, lead_full, kb * BK,
                                 n0 + (int)rank * (BN2 / 2));
                tma_load_2d_pair(base + 2 * kTileBytes + kHalfTileBytes, &map_b_lo, lead_full, kb * BK,
                                 n0 + (int)rank * (BN2 / 2));
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN2);
        uint32_t g = 0, it = 0;
        for (int t = pair; t < total_tiles; t += n_pairs_resident, ++it) {
            const uint32_t a = it & 1u, aph = (it >> 1) & 1u;
            mbar_wait(tempty0 + 8 * a, aph ^ 1u);  // both CTAs' epilogues drained this stage
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + a * BN2;
            for (int kb = 0; kb < kb_per; ++kb, ++g) {
                const uint32_t s = g % kStages2, ph = (g / kStages2) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = smem_u32(smem + s * kStageBytes2);
                const uint64_t a_hi = umma_desc_sw128(base);
                const uint64_t a_lo = umma_desc_sw128(base + kTileBytes);
                const uint64_t b_hi = umma_desc_sw128(base + 2 * kTileBytes);
                const uint64_t b_lo = umma_desc_sw128(base + 2 * kTileBytes + kHalfTileBytes);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                    umma2_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
                    umma2_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma2_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                umma2_commit_mc(empty0 + 8 * s, (uint16_t)0x3);
            }
            umma2_commit_mc(tfull0 + 8 * a, (uint16_t)0x3);  // accumulator stage a is complete
        }
    } else if (warp >= 2) {
        // ---- epilogue warps (both CTAs) ----
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        uint32_t it = 0;
        for (int t = pair; t < total_tiles; t += n_pairs_resident, ++it) {
            int m0, n0, item, a_col0;
            tile_coords(t, m0, n0, item, a_col0);
            const uint32_t a = it & 1u, aph = (it >> 1) & 1u;
            mbar_wait(tfull0 + 8 * a, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            gemm_epilogue<EPI, BN2>(tmem_base + a * BN2, m0, n0, q, lane,
                                    d_c64 ? d_c64 + (long long)item * pb.out_stride : nullptr, d_hi,
                                    d_lo, ldd, n_valid, scale);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty0 + 8 * a, 0));
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(512) : "memory");
}

#ifdef SOSAT_EXPERIMENTS  // cluster-of-four split-K kernel (measured slower, SOSAT_GEMM_QUAD=1)
// ---- two CTA pairs per tile: split-K inside a cluster of four ---------------------------------
// For sizes whose 256 x 256 pair tiles cannot fill the machine (1024^2: 32 tiles, 74 TPCs) a
// cluster of FOUR CTAs takes one tile: ranks (0,1) and (2,3) are two cta_group::2 pairs, each
// accumulating one half of K in its own TMEM.  Pair 1 then writes its fp32 sums to a scratch
// tile (plain coalesced 16-byte stores, L2-resident) and releases them to pair 0 through a
// cluster-scope mbarrier; pair 0 adds them in its epilogue.  No atomics, no extra kernels, and
// the K loop per CTA halves.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

template <int EPI>
__global__ void __launch_bounds__(128, 1)
gemm2k_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                     const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi,
                     const __grid_constant__ CUtensorMap map_b_lo, int num_kb, int n_tiles,
                     float4 *__restrict__ scratch, float *__restrict__ d_c64,
                     __nv_bfloat16 *__restrict__ d_hi, __nv_bfloat16 *__restrict__ d_lo, int ldd,
                     int n_valid, float scale) {
    constexpr int BN2 = 256;
    using Cfg = PairCfg<BN2>;
    constexpr int kStages2 = Cfg::kStages, kStageBytes2 = Cfg::kStageBytes,
                  kHalfTileBytes = Cfg::kHalfBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStages2 * kStageBytes2);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages2 + 2);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages2),
                   accum_bar = smem_u32(bars + 2 * kStages2),
                   partial_bar = smem_u32(bars + 2 * kStages2 + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();   // 0,1: pair of K slice 0; 2,3: pair of K slice 1
    const uint32_t prank = rank & 1u, lead = rank & ~1u, kslice = rank >> 1;
    const bool leader = prank == 0;
    const uint16_t pair_mask = (uint16_t)(0x3u << lead);
    const int tile = blockIdx.x >> 2;
    const int n0 = (tile % n_tiles) * BN2;
    const int m0 = (2 * (tile / n_tiles) + (int)prank) * BM;
    const int kb_per = num_kb / 2, kb0 = (int)kslice * kb_per;
    float4 *my_scratch = scratch + (size_t)(tile * 2 + (int)prank) * (BN2 / 4) * BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(full0 + 8 * s, 2);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum_bar, 1);
        mbar_init(partial_bar, 4);  // the four epilogue warps of the partner CTA of slice 1
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(BN2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();
    pdl_wait();
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kb_per; ++i) {
            const int kb = kb0 + i;
            const uint32_t s = i % kStages2, ph = (i / kStages2) & 1u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t lead_full = map_to_cta(full0 + 8 * s, lead);
            const uint32_t base = smem_u32(smem + s * kStageBytes2);
            if (leader) mbar_expect_tx(full0 + 8 * s, 2 * kStageBytes2);
            else mbar_arrive_cluster(lead_full);
            tma_load_2d_pair(base, &map_a_hi, lead_full, kb * BK, m0);
            tma_load_2d_pair(base + kTileBytes, &map_a_lo, lead_full, kb * BK, m0);
            tma_load_2d_pair(base + 2 * kTileBytes, &map_b_hi, lead_full, kb * BK,
                             n0 + (int)prank * (BN2 / 2));
            tma_load_2d_pair(base + 2 * kTileBytes + kHalfTileBytes, &map_b_lo, lead_full, kb * BK,
                             n0 + (int)prank * (BN2 / 2));
        }
    } else if (warp == 1 && lane == 0 && leader) {
        constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN2);
        for (int i = 0; i < kb_per; ++i) {
            const uint32_t s = i % kStages2, ph = (i / kStages2) & 1u;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(smem + s * kStageBytes2);
            const uint64_t a_hi = umma_desc_sw128(base);
            const uint64_t a_lo = umma_desc_sw128(base + kTileBytes);
            const uint64_t b_hi = umma_desc_sw128(base + 2 * kTileBytes);
            const uint64_t b_lo = umma_desc_sw128(base + 2 * kTileBytes + kHalfTileBytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                umma2_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (i | k) != 0);
                umma2_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                umma2_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
            }
            umma2_commit_mc(empty0 + 8 * s, pair_mask);
        }
        umma2_commit_mc(accum_bar, pair_mask);
    }
    __syncwarp();

    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (kslice == 1) {
        // K slice 1: hand the fp32 sums to the CTA of slice 0 that owns the same rows (rank - 2)
        store_partial_tile<BN2>(tmem_base, warp, lane, my_scratch);
        __threadfence();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_cta(partial_bar, rank - 2));
    } else {
        mbar_wait_cluster(partial_bar, 0);
        gemm_epilogue<EPI, BN2>(tmem_base, m0, n0, warp, lane, d_c64, d_hi, d_lo, ldd, n_valid, scale,
                                my_scratch);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(BN2) : "memory");
}

#endif  // SOSAT_EXPERIMENTS
// ---- operand preparation (HBM-bound elementwise kernels) -----------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Twiddle operand of BOTH passes; np = padded size, rows/cols beyond n are zero.
//   a[2k+c][c' np + m] = component (c, c') of W[k, m] folded as [Wr, -Wi; Wi, Wr]  (K planar)
// W[k, m] = exp(sign 2 pi i ((k - so)(m - si) mod n) / n): the shifts express the fftshift /
// ifftshift pairs of the reference (see dft_kind_params).  Pass 1 contracts it with the data
// operand X[m][c' np + n] over (c', n), pass 2 with D1[l][c' np + m] over (c', m): the same
// matrix in the same layout, so it is generated, stored and (second pass: mostly from L2) read
// once -- round 1 kept a K-interleaved copy for pass 1, 16 np^2 bytes more traffic per transform.
__global__ void twiddle_kernel(int n, int np, int so, int si, int sign, __nv_bfloat16 *a_hi,
                               __nv_bfloat16 *a_lo) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= np) return;
    float wr = 0.f, wi = 0.f;
    if (row < n && col < n) {
        long long j = ((long long)(row - so) * (long long)(col - si)) % n;
        if (j < 0) j += n;
        double s, c;
        sincospi((double)sign * 2.0 * (double)j / (double)n, &s, &c);  // exact integer phase
        wr = (float)c;
        wi = (float)s;
    }
    const float vals[2][2] = {{wr, -wi}, {wi, wr}};  // [c][c']
    const size_t ld = 2 * (size_t)np;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
            __nv_bfloat16 h, l;
            split_bf16(vals[c][cp], h, l);
            const size_t i = (size_t)(2 * row + c) * ld + (size_t)cp * np + col;
            a_hi[i] = h; a_lo[i] = l;
        }
}

// Twiddle operand of a ZERO-PADDED transform: the n_in x n_in field sits at rows / columns
// [pad, pad + n_in) of the n x n input, everything else is zero, so only those columns of W take
// part in either contraction:  a[2k+c][c' nip + j] = (c, c') component of W[k, pad + j], j < n_in
// (nip = n_in padded to the tile size; rows k >= n and columns j >= n_in are zero).
__global__ void twiddle_padded_kernel(int n, int np, int n_in, int nip, int pad, int so, int si,
                                      int sign, __nv_bfloat16 *a_hi, __nv_bfloat16 *a_lo) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= nip) return;
    float wr = 0.f, wi = 0.f;
    if (row < n && col < n_in) {
        long long j = ((long long)(row - so) * (long long)(col + pad - si)) % n;
        if (j < 0) j += n;
        double s, c;
        sincospi((double)sign * 2.0 * (double)j / (double)n, &s, &c);  // exact integer phase
        wr = (float)c;
        wi = (float)s;
    }
    const float vals[2][2] = {{wr, -wi}, {wi, wr}};  // [c][c']
    const size_t ld = 2 * (size_t)nip;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
            __nv_bfloat16 h, l;
            split_bf16(vals[c][cp], h, l);
            const size_t i = (size_t)(2 * row + c) * ld + (size_t)cp * nip + col;
            a_hi[i] = h; a_lo[i] = l;
        }
}

// Twiddle operands of the zoomed ("matrix DFT") transform on caller-chosen angles:
//   W1[l][n] = exp(-2 pi i y_n thy_l / lambda)   (pass 1, rows of b = y axis)
//   W2[k][m] = exp(-2 pi i x_m thx_k / lambda)   (pass 2, columns of b = x axis)
// coords: y[n], x[n] (n values each), angles: thy[n_out], thx[n_out]; the phase in cycles is
// reduced in fp64 before sincospi.  Same [c][c'] folding and K-planar layout as twiddle_kernel;
// the two passes have different matrices here, so both operand planes are used.
__global__ void twiddle_zoom_kernel(int n, int n_out, int np, const double *__restrict__ y,
                                    const double *__restrict__ thy, const double *__restrict__ x,
                                    const double *__restrict__ thx, double inv_lambda,
                                    __nv_bfloat16 *a1_hi, __nv_bfloat16 *a1_lo,
                                    __nv_bfloat16 *a2_hi, __nv_bfloat16 *a2_lo) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= np) return;
    float w1r = 0.f, w1i = 0.f, w2r = 0.f, w2i = 0.f;
    if (row < n_out && col < n) {
        double p = y[col] * thy[row] * inv_lambda, sn, cs;
        p -= rint(p);
        sincospi(-2.0 * p, &sn, &cs);
        w1r = (float)cs; w1i = (float)sn;
        p = x[col] * thx[row] * inv_lambda;
        p -= rint(p);
        sincospi(-2.0 * p, &sn, &cs);
        w2r = (float)cs; w2i = (float)sn;
    }
    const float v1[2][2] = {{w1r, -w1i}, {w1i, w1r}}, v2[2][2] = {{w2r, -w2i}, {w2i, w2r}};
    const size_t ld = 2 * (size_t)np;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
            __nv_bfloat16 h, l;
            const size_t i1 = (size_t)(2 * row + c) * ld + (size_t)cp * np + col;
            const size_t i2 = i1;
            split_bf16(v1[c][cp], h, l);
            a1_hi[i1] = h; a1_lo[i1] = l;
            split_bf16(v2[c][cp], h, l);
            a2_hi[i2] = h; a2_lo[i2] = l;
        }
}

// Data operand of pass 1: X[m][c' np + n] = component c' of b[n][m] (the transpose, zero padded
// to np, K planar like the twiddle operand), split into bf16 hi/lo.  32 x 32 tiles through shared
// memory: coalesced 8-byte reads along the rows of b, coalesced 2-byte writes along the rows of X.  LOAD(row, col) returns
// the complex sample b[row][col].
template <class Load>
__device__ __forceinline__ void prep_transposed(Load load, int n, int np, __nv_bfloat16 *b_hi,
                                                __nv_bfloat16 *b_lo) {
    __shared__ float2 tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;  // rows / columns of b
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + ty + 8 * q, c = c0 + tx;
        tile[ty + 8 * q][tx] = (r < n && c < n) ? load(r, c) : make_float2(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int m = c0 + ty + 8 * q, nn = r0 + tx;  // X row (column of b), X complex column
        const float2 v = tile[tx][ty + 8 * q];
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v.x, h0, l0);
        split_bf16(v.y, h1, l1);
        const size_t i = (size_t)m * 2 * np + nn;
        b_hi[i] = h0; b_hi[i + np] = h1;
        b_lo[i] = l0; b_lo[i + np] = l1;
    }
}

// blockIdx.z = batch item: input item z at b + z n^2, operand rows z np .. (z+1) np
__global__ void __launch_bounds__(256)
prep_b_c64_kernel(const float2 *__restrict__ b, int n, int np, __nv_bfloat16 *b_hi,
                  __nv_bfloat16 *b_lo) {
    pdl_wait();
    pdl_launch_dependents();
    const float2 *bz = b + (size_t)blockIdx.z * n * n;
    const size_t off = (size_t)blockIdx.z * np * 2 * np;
    prep_transposed([&](int r, int c) { return bz[(size_t)r * n + c]; }, n, np, b_hi + off,
                    b_lo + off);
}

// a2b / b2a input: |A| exp(i phi pi/180), opt_analyze.py:211, 224 (amplitude and degrees, float64)
__global__ void __launch_bounds__(256)
prep_b_a2b_kernel(const double *__restrict__ amp, const double *__restrict__ phi, int n, int np,
                  __nv_bfloat16 *b_hi, __nv_bfloat16 *b_lo) {
    pdl_wait();
    pdl_launch_dependents();
    prep_transposed(
        [&](int r, int c) {
            const size_t i = (size_t)r * n + c;
            double sn, cs;
            sincos(phi[i] * 3.141592653589793 / 180.0, &sn, &cs);
            const double a = fabs(amp[i]);
            return make_float2((float)(a * cs), (float)(a * sn));
        },
        n, np, b_hi, b_lo);
}

#ifdef SOSAT_EXPERIMENTS  // split-K helper kernels (measured slower, SOSAT_GEMM_SPLITK=1)
// Split-K pass 1 leaves fp32 sums in src [n]: split them into the bf16 hi/lo operand of pass 2,
// zero src in place for the next transform and zero the output the second pass reduces into.
__global__ void split_f32_kernel(float *__restrict__ src, long long n, __nv_bfloat16 *__restrict__ hi,
                                 __nv_bfloat16 *__restrict__ lo, float *__restrict__ zero_dst,
                                 long long n_zero) {
    pdl_wait();
    pdl_launch_dependents();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
        const float4 v = ((const float4 *)src)[i];
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]);
        split_bf16(v.y, h[1], l[1]);
        split_bf16(v.z, h[2], l[2]);
        split_bf16(v.w, h[3], l[3]);
        ((uint2 *)hi)[i] = make_uint2(
            (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
            (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
        ((uint2 *)lo)[i] = make_uint2(
            (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
            (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
        ((float4 *)src)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (zero_dst != src)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += stride)
            zero_dst[i] = 0.f;
}

__global__ void zero_f32_kernel(float *__restrict__ dst, long long n) {
    pdl_wait();
    pdl_launch_dependents();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4;
         i += (long long)gridDim.x * blockDim.x)
        ((float4 *)dst)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

#endif  // SOSAT_EXPERIMENTS
// a2b / b2a outputs from the complex64 transform: complex128 field and phase =
// arctan2(imag, real), opt_analyze.py:215-216, 228-229
__global__ void finish_a2b_kernel(const float2 *__restrict__ c64, long long n2,
                                  double2 *__restrict__ beam, double *__restrict__ phase) {
    pdl_wait();
    pdl_launch_dependents();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
         i += (long long)gridDim.x * blockDim.x) {
        const float2 v = c64[i];
        const double re = v.x, im = v.y;
        beam[i] = make_double2(re, im);
        phase[i] = atan2(im, re);
    }
}

// ---- host side -------------------------------------------------------------------------------
struct DftWorkspace {
    uint8_t *base;
    int n, np;
    __nv_bfloat16 *a1_hi, *a1_lo, *a2_hi, *a2_lo, *b1_hi, *b1_lo, *d1_hi, *d1_lo;
    float *c64;  // complex64 [n][n] staging for the float64 entry points (8 np^2 bytes)
};
static inline int pad128(int n) { return (n + 127) / 128 * 128; }

static size_t workspace_bytes(int n) {
    const size_t np = pad128(n);
    return 256 + 56 * np * np;
}

static DftWorkspace carve(void *ws, int n) {
    DftWorkspace w;
    w.base = (uint8_t *)ws;
    w.n = n;
    w.np = pad128(n);
    const size_t np2 = (size_t)w.np * w.np;
    uint8_t *p = w.base + 256;
    w.a1_hi = (__nv_bfloat16 *)p; p += 8 * np2;
    w.a1_lo = (__nv_bfloat16 *)p; p += 8 * np2;
    w.a2_hi = (__nv_bfloat16 *)p; p += 8 * np2;
    w.a2_lo = (__nv_bfloat16 *)p; p += 8 * np2;
    w.b1_hi = (__nv_bfloat16 *)p; p += 4 * np2;
    w.b1_lo = (__nv_bfloat16 *)p; p += 4 * np2;
    w.d1_hi = (__nv_bfloat16 *)p; p += 4 * np2;
    w.d1_lo = (__nv_bfloat16 *)p; p += 4 * np2;
    w.c64 = (float *)p;
    return w;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 row-major [rows][cols] with 128B-swizzled {BK x box_rows} boxes
static int make_map(CUtensorMap *m, const void *ptr, uint64_t rows, uint64_t cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(SOSAT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(SOSAT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// Twiddle operands and TMA descriptors depend only on (workspace, n, kind); they are built once
// per (device, workspace, n, kind) and remembered in a small host-side registry so that repeat
// transforms launch no extra work.  A caller that frees a planned workspace must call
// sosat_dft2_forget(ptr) before the address can be reused for another workspace.
struct PlanEntry {
    int device;
    const void *ws;
    int n, kind;
    bool wide;             // np % 256 == 0: both passes run as 2 x 2 clusters
    CUtensorMap maps[8];   // pass 1: A hi/lo, B hi/lo; pass 2: A hi/lo, B hi/lo
    CUtensorMap pair[8];   // the same operands with the boxes of the CTA-pair kernel (256 x 128)
    CUtensorMap pair256[8];  // ... and of its 256 x 256 form (128-row halves of the B tile)
};
// the one cached plan of the zero-padded batch transform (sosat_dft2_gemm_batch_padded)
struct PaddedPlan {
    int device;
    const void *ws;
    int n_in, pad, batch, kind;
    CUtensorMap m1[4], m2[4];
};
static PaddedPlan g_padded_plan = {-1, nullptr, 0, 0, 0, 0, {}, {}};
static PlanEntry g_plans[64];
static int g_num_plans = 0;
static std::mutex g_plan_mutex;  // the registry is shared by all host threads of the process

// The three centred transforms of opt_analyze.py as (output shift, input shift, sign, scale):
//   FORWARD   fftshift(fft2(fftshift(x)))      a2b :225-228, beam_convolve :76-82
//   INVERSE   ifftshift(ifft2(ifftshift(x)))   beam_convolve :86-88
//   B2A       fftshift(ifft2(x))               b2a :213-215 (the shifted input of :213 is
//                                              discarded by :214 -- quirk kept)
static int dft_kind_params(int kind, int n, int &so, int &si, int &sign, double &scale) {
    switch (kind) {
        case SOSAT_DFT_FORWARD: so = n / 2; si = (n + 1) / 2; sign = -1; scale = 1.0; return 0;
        case SOSAT_DFT_INVERSE: so = (n + 1) / 2; si = n / 2; sign = +1; scale = 1.0 / ((double)n * n); return 0;
        case SOSAT_DFT_B2A: so = n / 2; si = 0; sign = +1; scale = 1.0 / ((double)n * n); return 0;
    }
    return set_error(SOSAT_ERR_INVALID_ARG, "unknown transform kind %d", kind);
}

// Cluster shapes: pass 1 has grid (np/128, 2np/128), pass 2 (2np/128, np/128); the doubled
// dimension is always even, the other one only when np is a multiple of 256.
constexpr int kKindZoom = 100;  // twiddles written by the caller (sosat_dft2_zoom): never cached

// Returns a COPY of the plan (descriptors included): another host thread may re-plan the slot.
static int ensure_plan(const DftWorkspace &w, int kind, cudaStream_t st, PlanEntry *out) {
    int dev = 0, so = 0, si = 0, sign = 0;
    double scale;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    if (kind != kKindZoom)
        if (int rc = dft_kind_params(kind, w.n, so, si, sign, scale)) return rc;
    for (int i = 0; i < g_num_plans; ++i)
        if (g_plans[i].device == dev && g_plans[i].ws == w.base && g_plans[i].n == w.n &&
            g_plans[i].kind == kind && kind != kKindZoom) {
            *out = g_plans[i];
            return 0;
        }
    PlanEntry e;
    e.device = dev; e.ws = w.base; e.n = w.n; e.kind = kind;
    e.wide = (w.np % 256) == 0;
    const uint64_t np = w.np;
    const int half = 64, full = 128;
    int rc = 0;
    // pass 1: A = twiddle rows (l,c) (cluster y = M: always 2 -> B slices), B = X rows m
    //   A is shared along cluster x (CX = wide ? 2 : 1), B along cluster y (CY = 2).
    // The FFT kinds use ONE twiddle operand for both passes (the a2 planes); only the zoomed
    // transform, whose two passes have different matrices, fills a1 as well.
    const __nv_bfloat16 *t1_hi = kind == kKindZoom ? w.a1_hi : w.a2_hi;
    const __nv_bfloat16 *t1_lo = kind == kKindZoom ? w.a1_lo : w.a2_lo;
    if ((rc = make_map(&e.maps[0], t1_hi, 2 * np, 2 * np, e.wide ? half : full))) return rc;
    if ((rc = make_map(&e.maps[1], t1_lo, 2 * np, 2 * np, e.wide ? half : full))) return rc;
    if ((rc = make_map(&e.maps[2], w.b1_hi, np, 2 * np, half))) return rc;
    if ((rc = make_map(&e.maps[3], w.b1_lo, np, 2 * np, half))) return rc;
    // pass 2: A = pass-1 output [2 np][np] viewed as [np][2 np] rows l, B = twiddle a2 rows (k,c)
    //   CX = 2 always (A slices), CY = wide ? 2 : 1
    if ((rc = make_map(&e.maps[4], w.d1_hi, np, 2 * np, half))) return rc;
    if ((rc = make_map(&e.maps[5], w.d1_lo, np, 2 * np, half))) return rc;
    if ((rc = make_map(&e.maps[6], w.a2_hi, 2 * np, 2 * np, e.wide ? half : full))) return rc;
    if ((rc = make_map(&e.maps[7], w.a2_lo, 2 * np, 2 * np, e.wide ? half : full))) return rc;
    // CTA-pair kernel: whole 128-row A boxes, 64-row halves of the B tile
    if ((rc = make_map(&e.pair[0], t1_hi, 2 * np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair[1], t1_lo, 2 * np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair[2], w.b1_hi, np, 2 * np, half))) return rc;
    if ((rc = make_map(&e.pair[3], w.b1_lo, np, 2 * np, half))) return rc;
    if ((rc = make_map(&e.pair[4], w.d1_hi, np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair[5], w.d1_lo, np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair[6], w.a2_hi, 2 * np, 2 * np, half))) return rc;
    if ((rc = make_map(&e.pair[7], w.a2_lo, 2 * np, 2 * np, half))) return rc;
    for (int q = 0; q < 8; ++q) e.pair256[q] = e.pair[q];
    if ((rc = make_map(&e.pair256[2], w.b1_hi, np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair256[3], w.b1_lo, np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair256[6], w.a2_hi, 2 * np, 2 * np, full))) return rc;
    if ((rc = make_map(&e.pair256[7], w.a2_lo, 2 * np, 2 * np, full))) return rc;
    if (kind != kKindZoom) {
        dim3 grid((w.np + 127) / 128, w.np);
        twiddle_kernel<<<grid, 128, 0, st>>>(w.n, w.np, so, si, sign, w.a2_hi, w.a2_lo);
        SOSAT_CUDA_OK(cudaGetLastError());
    }
    int slot = -1;
    for (int i = 0; i < g_num_plans; ++i)
        if (g_plans[i].device == dev && g_plans[i].ws == w.base) slot = i;  // same buffer, re-planned
    if (slot < 0) slot = g_num_plans < 64 ? g_num_plans++ : 0;
    g_plans[slot] = e;
    *out = e;
    return 0;
}

// launch with programmatic stream serialization (the kernels call griddepcontrol.wait before
// touching global memory) and, for the GEMM passes, the cluster shape
template <class... KArgs, class... Args>
static int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                      int cx, int cy, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
    if (cx * cy > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = cx;
        attrs[na].val.clusterDim.y = cy;
        attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    SOSAT_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
    return 0;
}

template <int EPI, int CX, int CY>
static int launch_gemm(dim3 grid, cudaStream_t st, const CUtensorMap *m, int num_kb, float *c64,
                       __nv_bfloat16 *d_hi, __nv_bfloat16 *d_lo, int ldd, int n_valid, float scale) {
    static bool attr_set[64] = {false};  // per device: the attribute is per (function, context)
    int dev = 0;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        SOSAT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_kernel<EPI, CX, CY>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        attr_set[dev] = true;
    }
    return launch_pdl(gemm_bf16x3_kernel<EPI, CX, CY>, grid, dim3(128), kGemmSmem, st, CX, CY, m[0],
                      m[1], m[2], m[3], num_kb, c64, d_hi, d_lo, ldd, n_valid, scale);
}

#ifdef SOSAT_EXPERIMENTS  // launcher of the one-tile-per-CTA pair kernel
// grid = (M tiles of 128, N tiles of BN2): the pair kernel indexes M tiles with blockIdx.x
template <int EPI, int BN2>
static int launch_gemm_pair(dim3 grid, cudaStream_t st, const CUtensorMap *m, int num_kb, float *c64,
                            __nv_bfloat16 *d_hi, __nv_bfloat16 *d_lo, int ldd, int n_valid,
                            float scale, PairBatch pb = PairBatch{0, 0, 0, 0}) {
    static bool attr_set[64] = {false};
    int dev = 0;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        SOSAT_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16x3_kernel<EPI, BN2>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           PairCfg<BN2>::kSmem));
        attr_set[dev] = true;
    }
    return launch_pdl(gemm2_bf16x3_kernel<EPI, BN2>, grid, dim3(128), PairCfg<BN2>::kSmem, st, 2, 1,
                      m[0], m[1], m[2], m[3], num_kb, c64, d_hi, d_lo, ldd, n_valid, scale, pb);
}

#endif  // SOSAT_EXPERIMENTS
// persistent form of the 256 x 256 pair kernel: grid = one pair per TPC (or per tile if fewer)
template <int EPI>
static int launch_gemm_pair_persistent(int m_tiles, int n_tiles256, cudaStream_t st,
                                       const CUtensorMap *m, int num_kb, float *c64,
                                       __nv_bfloat16 *d_hi, __nv_bfloat16 *d_lo, int ldd, int n_valid,
                                       float scale, PairBatch pb = PairBatch{0, 0, 0, 0},
                                       int ksplit = 1) {
    static bool attr_set[64] = {false};
    int dev = 0;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    constexpr int smem = PairCfg<256>::kSmem + 64;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        SOSAT_CUDA_OK(cudaFuncSetAttribute(gemm2p_bf16x3_kernel<EPI>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    const int m_pairs = m_tiles / 2, tiles = m_pairs * n_tiles256 * ksplit;
    int pairs = num_sms() / 2;
    if (pairs > tiles) pairs = tiles;
    return launch_pdl(gemm2p_bf16x3_kernel<EPI>, dim3(2 * pairs), dim3(kPersistThreads), smem, st, 2, 1,
                      m[0], m[1], m[2], m[3], num_kb, m_pairs, n_tiles256, ksplit, c64, d_hi, d_lo,
                      ldd, n_valid, scale, pb);
}

// SOSAT_GEMM_PERSIST=0 selects the one-tile-per-CTA pair kernel for A/B runs
static bool use_persistent_pair() {
#ifdef SOSAT_EXPERIMENTS
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SOSAT_GEMM_PERSIST");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v != 0;
#else
    return true;
#endif
}

#ifdef SOSAT_EXPERIMENTS  // launchers / switches of the split-K experiments
template <int EPI>
static int launch_gemm_quad(int m_tiles, int n_tiles256, cudaStream_t st, const CUtensorMap *m,
                            int num_kb, float4 *scratch, float *c64, __nv_bfloat16 *d_hi,
                            __nv_bfloat16 *d_lo, int ldd, int n_valid, float scale) {
    static bool attr_set[64] = {false};
    int dev = 0;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    constexpr int smem = PairCfg<256>::kSmem + 64;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        SOSAT_CUDA_OK(cudaFuncSetAttribute(gemm2k_bf16x3_kernel<EPI>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    const int tiles = (m_tiles / 2) * n_tiles256;
    return launch_pdl(gemm2k_bf16x3_kernel<EPI>, dim3(4 * tiles), dim3(128), smem, st, 4, 1, m[0],
                      m[1], m[2], m[3], num_kb, n_tiles256, scratch, c64, d_hi, d_lo, ldd, n_valid,
                      scale);
}

// SOSAT_GEMM_QUAD=1 enables the cluster-of-four split-K kernel of the mid sizes (off by default:
// measured slower, see run_passes)
static bool use_quad_kernel() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SOSAT_GEMM_QUAD");
        v = e ? (atoi(e) != 0) : 0;
    }
    return v != 0;
}

// SOSAT_GEMM_SPLITK=1 enables the split-K path of the small sizes (off by default: measured
// slower, see run_passes)
static bool use_split_k() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SOSAT_GEMM_SPLITK");
        v = e ? (atoi(e) != 0) : 0;
    }
    return v != 0;
}

// Which UMMA kernel runs the passes.  Measured on B200 (1024^2 / 2048^2, ms per transform):
//   single-CTA UMMA, 2 x 2 multicast clusters, 128 x 128 tiles   0.071 / 0.353
//   CTA pair (cta_group::2), 256 x 128 pair tiles               0.089 / 0.527  (N = 128 UMMAs of a
//                                                                pair run at half rate)
//   CTA pair, 256 x 256 pair tiles                              0.098 / 0.324
// The 256 x 256 pair tile is the only shape that is not shared-memory-pipe bound, but it needs
// enough tiles to fill the machine: it is used when a pass has >= 120 CTAs (np >= 1536).
#endif  // SOSAT_EXPERIMENTS
// SOSAT_GEMM_PAIR = 0 / 128 / 256 overrides the choice for A/B runs.
static int pair_kernel_width(int np) {
#ifdef SOSAT_EXPERIMENTS
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("SOSAT_GEMM_PAIR");
        v = e ? atoi(e) : -1;
    }
    if (v >= 0) return v;
#endif
    return (np % 256 == 0 && (2 * np / BM) * (np / 256) >= 120) ? 256 : 0;
}

// the two GEMM passes; b1 must already hold the split, transposed data operand.  The complex64
// result (times scale) is written to d_c64 [n][n].
static int run_passes(const DftWorkspace &w, const PlanEntry &pl, float scale, float *d_c64,
                      cudaStream_t st, int n_out = 0) {
    if (n_out <= 0) n_out = w.n;  // output is n_out x n_out complex64
    const int num_kb = 2 * w.np / BK;
    const dim3 g1(w.np / BN, 2 * w.np / BM), g2(2 * w.np / BN, w.np / BM);
    int rc;
    const int pw = BK == 64 ? pair_kernel_width(w.np) : 0;
    const unsigned mt1 = 2 * w.np / BM, mt2 = w.np / BM;  // M tiles of pass 1 / pass 2
#ifdef SOSAT_EXPERIMENTS
    // Cluster-of-four split-K (two CTA pairs per pair tile, partial sums through an L2-resident
    // scratch tile) for sizes with at most one tile per four SMs: 512^2 ... 1024^2.  Scratch:
    // pass 1 uses the complex64 staging area, pass 2 the (by then consumed) X operand.
    // Parity-green but OFF by default: a pair-tile CTA with half the K loop (16 k-blocks) still
    // takes 27-28 us at 1024^2 -- ~9 us of prologue (TMEM allocation, four-CTA cluster sync, first
    // fills) and epilogue around a mainloop at ~78 % of the UMMA rate -- i.e. as long as the
    // single-CTA kernel with the full K loop: 0.083 vs 0.072 ms.
    if (BK == 64 && pl.wide && pw == 0 && use_quad_kernel() && !use_split_k() &&
        (int)(mt1 / 2) * (w.np / 256) * 4 <= num_sms() && num_kb % 2 == 0 && num_kb / 2 >= 4) {
        if ((rc = launch_gemm_quad<EPI_SPLIT_BF16>((int)mt1, w.np / 256, st, pl.pair256, num_kb,
                                                   (float4 *)w.c64, nullptr, w.d1_hi, w.d1_lo, w.np,
                                                   0, 1.f))) return rc;
        return launch_gemm_quad<EPI_C64>((int)mt2, 2 * w.np / 256, st, pl.pair256 + 4, num_kb,
                                         (float4 *)w.b1_hi, d_c64, nullptr, nullptr, 0, n_out, scale);
    }
    // Split-K on the pair kernel for sizes whose pair tiles alone cannot fill the machine
    // (1024^2: 32 pair tiles on 74 TPCs): each tile's K range is cut into ksplit slices that run
    // on different pairs and meet through vector red.global.add in a zeroed fp32 output.
    //   prep -> zero D1f -> pass 1 (red into D1f [2np][np] fp32, the c64 staging area)
    //   -> split (D1f -> bf16 hi/lo, zeroes D1f and the output) -> pass 2 (red into the output)
    // Parity-green but OFF by default: at 1024^2 each split-K pass takes 28.4 us (the mainloop
    // halves to ~13 us, but 1M float4 reductions per pass through L2 cost what it saves) against
    // 27 us for the single-CTA kernel, plus 11 us of zero/split kernels: 0.084 vs 0.072 ms.
    if (BK == 64 && pl.wide && use_persistent_pair() && pw == 0 && use_split_k()) {
        const int tiles = (int)(mt1 / 2) * (w.np / 256);
        int ksplit = 1;
        while (ksplit < 8 && tiles * ksplit * 2 <= num_sms() / 2 && num_kb / (ksplit * 2) >= 8 &&
               num_kb % (ksplit * 2) == 0)
            ksplit *= 2;
        if (ksplit > 1) {
            float *d1f = w.c64;  // 2 np x np floats
            const long long n1 = 2LL * w.np * w.np, n_out2 = 2LL * n_out * n_out;
            long long blocks = (n1 / 4 + 255) / 256;
            if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
            if ((rc = launch_pdl(zero_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, 1, 1, d1f,
                                 n1))) return rc;
            if ((rc = launch_gemm_pair_persistent<EPI_RED_F32>((int)mt1, w.np / 256, st, pl.pair256,
                                                               num_kb, d1f, nullptr, nullptr, w.np, 0,
                                                               1.f, PairBatch{0, 0, 0, 0}, ksplit)))
                return rc;
            if ((rc = launch_pdl(split_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, 1, 1, d1f,
                                 n1, w.d1_hi, w.d1_lo, d_c64, n_out2))) return rc;
            return launch_gemm_pair_persistent<EPI_RED_C64>((int)mt2, 2 * w.np / 256, st,
                                                            pl.pair256 + 4, num_kb, d_c64, nullptr,
                                                            nullptr, 0, n_out, scale,
                                                            PairBatch{0, 0, 0, 0}, ksplit);
        }
    }
#endif  // SOSAT_EXPERIMENTS
    if (pw == 256 && pl.wide && use_persistent_pair()) {
        if ((rc = launch_gemm_pair_persistent<EPI_SPLIT_BF16>((int)mt1, w.np / 256, st, pl.pair256,
                                                              num_kb, nullptr, w.d1_hi, w.d1_lo,
                                                              w.np, 0, 1.f))) return rc;
        return launch_gemm_pair_persistent<EPI_C64>((int)mt2, 2 * w.np / 256, st, pl.pair256 + 4,
                                                    num_kb, d_c64, nullptr, nullptr, 0, n_out, scale);
    }
#ifdef SOSAT_EXPERIMENTS
    if (pw == 256 && pl.wide) {
        if ((rc = launch_gemm_pair<EPI_SPLIT_BF16, 256>(dim3(mt1, w.np / 256), st, pl.pair256, num_kb,
                                                        nullptr, w.d1_hi, w.d1_lo, w.np, 0, 1.f)))
            return rc;
        return launch_gemm_pair<EPI_C64, 256>(dim3(mt2, 2 * w.np / 256), st, pl.pair256 + 4, num_kb,
                                              d_c64, nullptr, nullptr, 0, n_out, scale);
    }
    if (pw == 128) {
        // pass 1 has an even number of M tiles for every size; pass 2 only when np % 256 == 0
        if ((rc = launch_gemm_pair<EPI_SPLIT_BF16, 128>(dim3(mt1, w.np / BN), st, pl.pair, num_kb,
                                                        nullptr, w.d1_hi, w.d1_lo, w.np, 0, 1.f)))
            return rc;
        if (pl.wide)
            return launch_gemm_pair<EPI_C64, 128>(dim3(mt2, 2 * w.np / BN), st, pl.pair + 4, num_kb,
                                                  d_c64, nullptr, nullptr, 0, n_out, scale);
        return launch_gemm<EPI_C64, 2, 1>(g2, st, pl.maps + 4, num_kb, d_c64, nullptr, nullptr, 0,
                                          n_out, scale);
    }
#endif  // SOSAT_EXPERIMENTS
    if (pl.wide) {
        if ((rc = launch_gemm<EPI_SPLIT_BF16, 2, 2>(g1, st, pl.maps, num_kb, nullptr, w.d1_hi,
                                                    w.d1_lo, w.np, 0, 1.f))) return rc;
        return launch_gemm<EPI_C64, 2, 2>(g2, st, pl.maps + 4, num_kb, d_c64, nullptr, nullptr, 0,
                                          n_out, scale);
    }
    if ((rc = launch_gemm<EPI_SPLIT_BF16, 1, 2>(g1, st, pl.maps, num_kb, nullptr, w.d1_hi, w.d1_lo,
                                                w.np, 0, 1.f))) return rc;
    return launch_gemm<EPI_C64, 2, 1>(g2, st, pl.maps + 4, num_kb, d_c64, nullptr, nullptr, 0, n_out,
                                      scale);
}

static int check_ws(const void *ws, size_t bytes, int n) {
    SOSAT_REQUIRE(n >= 1 && n <= 16384, "transform size n=%d out of range [1, 16384]", n);
    SOSAT_REQUIRE(ws != nullptr, "null workspace");
    SOSAT_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
    SOSAT_REQUIRE(bytes >= workspace_bytes(n), "workspace too small: %zu < %zu bytes", bytes,
                  workspace_bytes(n));
    return 0;
}

}  // namespace sosat

using namespace sosat;

extern "C" size_t sosat_dft2_workspace_bytes(int n) { return n >= 1 ? workspace_bytes(n) : 0; }

extern "C" int sosat_dft2_forget(const void *d_workspace) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    for (int i = 0; i < g_num_plans; ++i)
        if (g_plans[i].ws == d_workspace) g_plans[i] = g_plans[--g_num_plans], --i;
    if (g_padded_plan.ws == d_workspace) g_padded_plan.ws = nullptr;
    return SOSAT_OK;
}

extern "C" int sosat_dft2_gemm_kind(const float *d_b, int n, float *d_B, int kind,
                                    void *d_workspace, size_t workspace_bytes_, void *stream) {
    SOSAT_REQUIRE(d_b && d_B, "sosat_dft2_gemm: null pointer");
    if (int rc = check_ws(d_workspace, workspace_bytes_, n)) return rc;
    int so, si, sign;
    double scale;
    if (int rc = dft_kind_params(kind, n, so, si, sign, scale)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DftWorkspace w = carve(d_workspace, n);
    PlanEntry pl;
    if (int rc = ensure_plan(w, kind, st, &pl)) return rc;
    const dim3 grid(w.np / 32, w.np / 32);
    if (int rc = launch_pdl(prep_b_c64_kernel, grid, dim3(32, 8), 0, st, 1, 1, (const float2 *)d_b, n,
                            w.np, w.b1_hi, w.b1_lo)) return rc;
    return run_passes(w, pl, (float)scale, d_B, st);
}

extern "C" int sosat_dft2_zoom(const float *d_b, int n, float *d_B, int n_out, const double *h_y,
                               const double *h_x, const double *h_thy, const double *h_thx,
                               double inv_lambda, void *d_workspace, size_t workspace_bytes_,
                               void *stream) {
    SOSAT_REQUIRE(d_b && d_B && h_y && h_x && h_thy && h_thx, "sosat_dft2_zoom: null pointer");
    if (int rc = check_ws(d_workspace, workspace_bytes_, n)) return rc;
    SOSAT_REQUIRE(n_out >= 1 && n_out <= pad128(n), "n_out=%d must be in [1, %d] for n=%d", n_out,
                  pad128(n), n);
    cudaStream_t st = (cudaStream_t)stream;
    DftWorkspace w = carve(d_workspace, n);
    PlanEntry pl;
    if (int rc = ensure_plan(w, kKindZoom, st, &pl)) return rc;
    // coordinates and angles staged at the start of the (otherwise unused here) complex64 area
    double *d_y = (double *)w.c64, *d_x = d_y + w.np, *d_thy = d_x + w.np, *d_thx = d_thy + w.np;
    SOSAT_CUDA_OK(cudaMemcpyAsync(d_y, h_y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    SOSAT_CUDA_OK(cudaMemcpyAsync(d_x, h_x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    SOSAT_CUDA_OK(cudaMemcpyAsync(d_thy, h_thy, sizeof(double) * n_out, cudaMemcpyHostToDevice, st));
    SOSAT_CUDA_OK(cudaMemcpyAsync(d_thx, h_thx, sizeof(double) * n_out, cudaMemcpyHostToDevice, st));
    const dim3 tg((w.np + 127) / 128, w.np);
    twiddle_zoom_kernel<<<tg, 128, 0, st>>>(n, n_out, w.np, d_y, d_thy, d_x, d_thx, inv_lambda,
                                            w.a1_hi, w.a1_lo, w.a2_hi, w.a2_lo);
    SOSAT_CUDA_OK(cudaGetLastError());
    const dim3 grid(w.np / 32, w.np / 32);
    if (int rc = launch_pdl(prep_b_c64_kernel, grid, dim3(32, 8), 0, st, 1, 1, (const float2 *)d_b, n,
                            w.np, w.b1_hi, w.b1_lo)) return rc;
    return run_passes(w, pl, 1.0f, d_B, st, n_out);
}

// ---- batched transform: many fields of one size through ONE pair of UMMA launches -------------
static size_t batch_workspace_bytes(int n, int batch) {
    const size_t np = pad128(n);
    const size_t need = 256 + 32 * np * np + 16 * np * np * (size_t)batch;
    return need > workspace_bytes(n) ? need : workspace_bytes(n);
}

extern "C" size_t sosat_dft2_batch_workspace_bytes(int n, int batch) {
    return (n >= 1 && batch >= 1) ? batch_workspace_bytes(n, batch) : 0;
}

extern "C" int sosat_dft2_gemm_batch(const float *d_b, int n, int batch, float *d_B, int kind,
                                     void *d_workspace, size_t workspace_bytes_, void *stream) {
    SOSAT_REQUIRE(d_b && d_B, "sosat_dft2_gemm_batch: null pointer");
    SOSAT_REQUIRE(batch >= 1 && batch <= 4096, "batch=%d out of range [1, 4096]", batch);
    if (int rc = check_ws(d_workspace, workspace_bytes_, n)) return rc;
    SOSAT_REQUIRE(workspace_bytes_ >= batch_workspace_bytes(n, batch),
                  "workspace too small for batch %d: %zu < %zu bytes", batch, workspace_bytes_,
                  batch_workspace_bytes(n, batch));
    int so, si, sign;
    double scale;
    if (int rc = dft_kind_params(kind, n, so, si, sign, scale)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DftWorkspace w = carve(d_workspace, n);
    PlanEntry pl;
    if (int rc = ensure_plan(w, kind, st, &pl)) return rc;
    const size_t np = w.np, n2 = (size_t)n * n;
    if (!(pl.wide && BK == 64)) {
        // sizes the pair kernel does not take (np not a multiple of 256): one transform at a time
        const dim3 grid(w.np / 32, w.np / 32);
        for (int i = 0; i < batch; ++i) {
            if (int rc = launch_pdl(prep_b_c64_kernel, grid, dim3(32, 8), 0, st, 1, 1,
                                    (const float2 *)d_b + i * n2, n, w.np, w.b1_hi, w.b1_lo)) return rc;
            if (int rc = run_passes(w, pl, (float)scale, d_B + 2 * i * n2, st)) return rc;
        }
        return SOSAT_OK;
    }
    // operands of the whole batch behind the twiddles: X [batch np][2 np], D1 [2 np][batch np]
    __nv_bfloat16 *xb_hi = w.b1_hi, *xb_lo = xb_hi + 2 * np * np * batch;
    __nv_bfloat16 *d1_hi = xb_lo + 2 * np * np * batch, *d1_lo = d1_hi + 2 * np * np * batch;
    CUtensorMap m1[4], m2[4];
    int rc = 0;
    m1[0] = pl.pair256[0];
    m1[1] = pl.pair256[1];
    if ((rc = make_map(&m1[2], xb_hi, np * batch, 2 * np, 128))) return rc;
    if ((rc = make_map(&m1[3], xb_lo, np * batch, 2 * np, 128))) return rc;
    if ((rc = make_map(&m2[0], d1_hi, np, 2 * np * batch, 128))) return rc;
    if ((rc = make_map(&m2[1], d1_lo, np, 2 * np * batch, 128))) return rc;
    m2[2] = pl.pair256[6];
    m2[3] = pl.pair256[7];
    const dim3 pgrid(w.np / 32, w.np / 32, batch);
    if ((rc = launch_pdl(prep_b_c64_kernel, pgrid, dim3(32, 8), 0, st, 1, 1, (const float2 *)d_b, n,
                         w.np, xb_hi, xb_lo))) return rc;
    const int num_kb = 2 * w.np / BK;
    const PairBatch pb{w.np / BK, w.np * batch, w.np / BM, (long long)(2 * n2)};
    if (use_persistent_pair()) {
        if ((rc = launch_gemm_pair_persistent<EPI_SPLIT_BF16>(2 * w.np / BM, w.np * batch / 256, st, m1,
                                                              num_kb, nullptr, d1_hi, d1_lo,
                                                              w.np * batch, 0, 1.f))) return rc;
        return launch_gemm_pair_persistent<EPI_C64>(w.np / BM * batch, 2 * w.np / 256, st, m2, num_kb,
                                                    d_B, nullptr, nullptr, 0, w.n, (float)scale, pb);
    }
#ifdef SOSAT_EXPERIMENTS
    if ((rc = launch_gemm_pair<EPI_SPLIT_BF16, 256>(dim3(2 * w.np / BM, w.np * batch / 256), st, m1,
                                                    num_kb, nullptr, d1_hi, d1_lo, w.np * batch, 0,
                                                    1.f))) return rc;
    return launch_gemm_pair<EPI_C64, 256>(dim3(w.np / BM * batch, 2 * w.np / 256), st, m2, num_kb, d_B,
                                          nullptr, nullptr, 0, w.n, (float)scale, pb);
#else
    return set_error(SOSAT_ERR_UNSUPPORTED, "unreachable: the persistent pair kernel is the only batch path");
#endif
}

// ---- batched transform of ZERO-PADDED fields -------------------------------------------------
// A sweep zero-pads its near fields before the transform for a finer angle sampling (zero_pad,
// opt_analyze.py:20-32; beam_sweep's `pad`).  The zeros need not be multiplied: with the field at
// rows / columns [pad, pad + n_in) only n_in columns of W take part in either contraction, so both
// passes run with K = 2 n_in instead of 2 n and pass 1 produces n_in instead of n columns -- the
// same two UMMA launches on compact operands (twiddle_padded_kernel), 3/8 of the work at
// n = 2 n_in, and the padded copy of the input is never materialised.
static inline size_t pad256(size_t n) { return (n + 255) / 256 * 256; }

static size_t padded_workspace_bytes(int n_in, int pad, int batch) {
    const size_t np = pad128(n_in + 2 * pad), nip = pad256(n_in);
    // twiddle [2 np][2 nip] hi/lo, X [batch nip][2 nip] hi/lo, D1 [2 np][batch nip] hi/lo
    return 256 + 2 * (2 * np * 2 * nip * 2) + 2 * ((size_t)batch * nip * 2 * nip * 2) +
           2 * (2 * np * (size_t)batch * nip * 2);
}

extern "C" size_t sosat_dft2_batch_padded_workspace_bytes(int n_in, int pad, int batch) {
    return (n_in >= 1 && pad >= 0 && batch >= 1) ? padded_workspace_bytes(n_in, pad, batch) : 0;
}

extern "C" int sosat_dft2_gemm_batch_padded(const float *d_b, int n_in, int pad, int batch, float *d_B,
                                            int kind, void *d_workspace, size_t workspace_bytes_,
                                            void *stream) {
    SOSAT_REQUIRE(d_b && d_B && d_workspace, "sosat_dft2_gemm_batch_padded: null pointer");
    SOSAT_REQUIRE(n_in >= 1 && pad >= 0 && n_in + 2 * pad <= 16384, "n_in=%d pad=%d out of range", n_in, pad);
    SOSAT_REQUIRE(batch >= 1 && batch <= 4096, "batch=%d out of range [1, 4096]", batch);
    SOSAT_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
    const int n = n_in + 2 * pad, np = pad128(n), nip = (int)pad256(n_in);
    SOSAT_REQUIRE(np % 256 == 0 && BK == 64,
                  "the padded batch transform needs a padded size that is a multiple of 256 (n = %d)", n);
    SOSAT_REQUIRE(workspace_bytes_ >= padded_workspace_bytes(n_in, pad, batch),
                  "workspace too small: %zu < %zu bytes", workspace_bytes_,
                  padded_workspace_bytes(n_in, pad, batch));
    int so, si, sign, dev = 0;
    double scale;
    if (int rc = dft_kind_params(kind, n, so, si, sign, scale)) return rc;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *p = (uint8_t *)d_workspace + 256;
    const size_t tw = (size_t)2 * np * 2 * nip, xs = (size_t)batch * nip * 2 * nip,
                 ds = (size_t)2 * np * batch * nip;
    __nv_bfloat16 *t_hi = (__nv_bfloat16 *)p, *t_lo = t_hi + tw, *x_hi = t_lo + tw, *x_lo = x_hi + xs,
                  *d1_hi = x_lo + xs, *d1_lo = d1_hi + ds;
    PaddedPlan pl;
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        PaddedPlan &g = g_padded_plan;
        if (!(g.device == dev && g.ws == d_workspace && g.n_in == n_in && g.pad == pad &&
              g.batch == batch && g.kind == kind)) {
            int rc = 0;
            // pass 1: A = compact twiddle rows (l,c), B = X rows (item, m); pass 2: A = D1 viewed
            // [np][2 nip batch] rows l, B = the same twiddle rows (k,c)
            if ((rc = make_map(&g.m1[0], t_hi, 2 * (uint64_t)np, 2 * (uint64_t)nip, 128))) return rc;
            if ((rc = make_map(&g.m1[1], t_lo, 2 * (uint64_t)np, 2 * (uint64_t)nip, 128))) return rc;
            if ((rc = make_map(&g.m1[2], x_hi, (uint64_t)batch * nip, 2 * (uint64_t)nip, 128))) return rc;
            if ((rc = make_map(&g.m1[3], x_lo, (uint64_t)batch * nip, 2 * (uint64_t)nip, 128))) return rc;
            if ((rc = make_map(&g.m2[0], d1_hi, (uint64_t)np, 2 * (uint64_t)nip * batch, 128))) return rc;
            if ((rc = make_map(&g.m2[1], d1_lo, (uint64_t)np, 2 * (uint64_t)nip * batch, 128))) return rc;
            g.m2[2] = g.m1[0];
            g.m2[3] = g.m1[1];
            const dim3 grid((nip + 127) / 128, np);
            twiddle_padded_kernel<<<grid, 128, 0, st>>>(n, np, n_in, nip, pad, so, si, sign, t_hi, t_lo);
            SOSAT_CUDA_OK(cudaGetLastError());
            g.device = dev; g.ws = d_workspace; g.n_in = n_in; g.pad = pad; g.batch = batch; g.kind = kind;
        }
        pl = g;
    }
    int rc = 0;
    const dim3 pgrid(nip / 32, nip / 32, batch);
    if ((rc = launch_pdl(prep_b_c64_kernel, pgrid, dim3(32, 8), 0, st, 1, 1, (const float2 *)d_b, n_in,
                         nip, x_hi, x_lo))) return rc;
    const int num_kb = 2 * nip / BK;
    const PairBatch pb{nip / BK, nip * batch, np / BM, (long long)(2 * (size_t)n * n)};
    if ((rc = launch_gemm_pair_persistent<EPI_SPLIT_BF16>(2 * np / BM, nip * batch / 256, st, pl.m1, num_kb,
                                                          nullptr, d1_hi, d1_lo, nip * batch, 0, 1.f)))
        return rc;
    return launch_gemm_pair_persistent<EPI_C64>(np / BM * batch, 2 * np / 256, st, pl.m2, num_kb, d_B,
                                                nullptr, nullptr, 0, n, (float)scale, pb);
}

extern "C" int sosat_dft2_gemm(const float *d_b, int n, float *d_B, void *d_workspace,
                               size_t workspace_bytes_, void *stream) {
    return sosat_dft2_gemm_kind(d_b, n, d_B, SOSAT_DFT_FORWARD, d_workspace, workspace_bytes_,
                                stream);
}

static int amp_phase_transform(const double *d_amp, const double *d_phi_deg, int n, int kind,
                               double *d_out, double *d_phase, void *d_workspace,
                               size_t workspace_bytes_, void *stream) {
    SOSAT_REQUIRE(d_amp && d_phi_deg && d_out && d_phase, "null pointer");
    if (int rc = check_ws(d_workspace, workspace_bytes_, n)) return rc;
    int so, si, sign;
    double scale;
    if (int rc = dft_kind_params(kind, n, so, si, sign, scale)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    DftWorkspace w = carve(d_workspace, n);
    PlanEntry pl;
    if (int rc = ensure_plan(w, kind, st, &pl)) return rc;
    const dim3 grid(w.np / 32, w.np / 32);
    if (int rc = launch_pdl(prep_b_a2b_kernel, grid, dim3(32, 8), 0, st, 1, 1, d_amp, d_phi_deg, n,
                            w.np, w.b1_hi, w.b1_lo)) return rc;
    if (int rc = run_passes(w, pl, (float)scale, w.c64, st)) return rc;  // complex64 [n][n]
    const long long n2 = (long long)n * n;
    long long blocks = (n2 + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    return launch_pdl(finish_a2b_kernel, dim3((unsigned)blocks), dim3(256), 0, st, 1, 1,
                      (const float2 *)w.c64, n2, (double2 *)d_out, d_phase);
}

extern "C" int sosat_a2b(const double *d_apert, const double *d_phi_deg, int n, double *d_beam,
                         double *d_phase, void *d_workspace, size_t workspace_bytes_,
                         void *stream) {
    return amp_phase_transform(d_apert, d_phi_deg, n, SOSAT_DFT_FORWARD, d_beam, d_phase,
                               d_workspace, workspace_bytes_, stream);
}

extern "C" int sosat_b2a(const double *d_beam, const double *d_phase_deg, int n, double *d_aper,
                         double *d_aper_phase, void *d_workspace, size_t workspace_bytes_,
                         void *stream) {
    return amp_phase_transform(d_beam, d_phase_deg, n, SOSAT_DFT_B2A, d_aper, d_aper_phase,
                               d_workspace, workspace_bytes_, stream);
}

extern "C" int sosat_a2b_host(const double *h_apert, const double *h_phi_deg, int n,
                              double *h_beam, double *h_phase) {
    SOSAT_REQUIRE(h_apert && h_phi_deg && h_beam && h_phase, "sosat_a2b_host: null pointer");
    SOSAT_REQUIRE(n >= 1 && n <= 16384, "transform size n=%d out of range", n);
    const size_t n2 = (size_t)n * n, wsb = workspace_bytes(n);
    uint8_t *buf = nullptr;
    // one allocation: workspace | apert | phi | beam | phase
    SOSAT_CUDA_OK(cudaMalloc(&buf, wsb + n2 * 8 * 5));
    double *d_ap = (double *)(buf + wsb), *d_ph = d_ap + n2, *d_beam = d_ph + n2,
           *d_phase = d_beam + 2 * n2;
    int rc = 0;
    sosat_dft2_forget(buf);
    cudaError_t e = cudaMemcpy(d_ap, h_apert, n2 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_ph, h_phi_deg, n2 * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = set_error(SOSAT_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
    if (!rc) rc = sosat_a2b(d_ap, d_ph, n, d_beam, d_phase, buf, wsb, nullptr);
    if (!rc) {
        e = cudaMemcpy(h_beam, d_beam, n2 * 16, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(h_phase, d_phase, n2 * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_error(SOSAT_ERR_CUDA, "D2H copy: %s", cudaGetErrorString(e));
    }
    sosat_dft2_forget(buf);
    cudaFree(buf);
    return rc;
}

// ---- beam_convolve helpers (opt_analyze.py:60-88) ---------------------------------------------
namespace sosat {
// cos-tapered rectangular horn aperture "disc" (opt_analyze.py:65-73) as complex64
__global__ void horn_aperture_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                     long long n, double apert1, double apert2,
                                     float2 *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double xv = x[i], yv = y[i];
        double d = apert1 < apert2 ? cospi(yv / apert2) : cospi(xv / apert1);
        if (!(fabs(xv) <= apert1 / 2) || !(fabs(yv) <= apert2 / 2)) d = 0.0;
        out[i] = make_float2((float)d, 0.f);
    }
}
__global__ void cmul_kernel(const float2 *__restrict__ a, const float2 *__restrict__ b, long long n,
                            float2 *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float2 u = a[i], v = b[i];
        out[i] = make_float2(u.x * v.x - u.y * v.y, u.x * v.y + u.y * v.x);
    }
}
}  // namespace sosat

extern "C" int sosat_horn_aperture(const double *d_x, const double *d_y, int64_t n, double apert1,
                                   double apert2, float *d_out, void *stream) {
    SOSAT_REQUIRE(d_x && d_y && d_out, "sosat_horn_aperture: null pointer");
    SOSAT_REQUIRE(n >= 0, "negative size");
    if (n == 0) return SOSAT_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    horn_aperture_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_x, d_y, n, apert1, apert2, (float2 *)d_out);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_cmul(const float *d_a, const float *d_b, int64_t n, float *d_out,
                          void *stream) {
    SOSAT_REQUIRE(d_a && d_b && d_out, "sosat_cmul: null pointer");
    SOSAT_REQUIRE(n >= 0, "negative size");
    if (n == 0) return SOSAT_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    cmul_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float2 *)d_a, (const float2 *)d_b, n, (float2 *)d_out);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}
