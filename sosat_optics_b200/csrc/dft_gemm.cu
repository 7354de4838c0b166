// Stage 2 of the hot path on sm_100a: the far-field transform of opt_analyze.a2b
// (opt_analyze.py:220-230),  B = fftshift(fft2(fftshift(b)))  ==  W b W^T  with
//     W[k, n] = exp(-2 pi i ((k - floor(N/2)) (n - ceil(N/2)) mod N) / N),
// executed as a dense separable complex contraction: two real GEMM passes on tcgen05 tensor
// cores (UTCHMMA), operands staged by TMA into 128B-swizzled shared memory, fp32 accumulators
// in TMEM.  A single bf16 pass is not accurate enough for the 1e-4-of-peak tolerance
// (SURVEY section 7), so every operand is split x = hi + lo (two bf16) and each k-block issues
// the three products hi*hi + hi*lo + lo*hi into the same accumulator ("bf16x3").
//
//   pass 1:  D1[(l,c), m] = sum_n  Wc[(l,c), (n,c')] * b[m, (n,c')]        (column transform)
//   pass 2:  D2[(k,c), l] = sum_m  Wr[(k,c), (c',m)] * D1[l, (c',m)]       (row transform)
//
// (.,c) = real/imag component.  The complex structure is folded into the twiddle operand:
// row (l,re) = [Wr, -Wi], row (l,im) = [Wi, Wr] along K, so the data operand is just the
// complex array viewed as reals.  Both passes are "TN" GEMMs with K-major operands; pass 1
// writes its result already split into bf16 hi/lo in exactly the layout pass 2 reads.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>

#include "common.cuh"

namespace sosat {

constexpr int BM = 128, BN = 128, BK = 64;  // CTA tile; BK * 2 B = one 128 B swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 3;
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

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row swizzle atoms 1024 B apart
// (SBO = 1024 B; LBO unused for swizzled K-major layouts and set to 1), descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

enum { EPI_F32 = 0, EPI_SPLIT_BF16 = 1 };

// D[M][N] (+)= A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T ; one 128 x 128 tile per CTA.
// warp 0 lane 0: TMA producer.  warp 1 lane 0: MMA issuer.  all 4 warps: epilogue.
template <int EPI>
__global__ void __launch_bounds__(128, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                   const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi,
                   const __grid_constant__ CUtensorMap map_b_lo, int num_kb, float *__restrict__ d_f32,
                   __nv_bfloat16 *__restrict__ d_hi, __nv_bfloat16 *__restrict__ d_lo, int ldd) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kStages * kStageBytes);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * kStages + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages),
                   accum_bar = smem_u32(bars + 2 * kStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
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

    if (warp == 0 && lane == 0) {
        // ---- TMA producer ----
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t full = full0 + 8 * s;
            const uint32_t base = smem_u32(smem + s * kStageBytes);
            mbar_expect_tx(full, kStageBytes);
            tma_load_2d(base + 0 * kTileBytes, &map_a_hi, full, kb * BK, m0);
            tma_load_2d(base + 1 * kTileBytes, &map_a_lo, full, kb * BK, m0);
            tma_load_2d(base + 2 * kTileBytes, &map_b_hi, full, kb * BK, n0);
            tma_load_2d(base + 3 * kTileBytes, &map_b_lo, full, kb * BK, n0);
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
            umma_commit(empty0 + 8 * s);  // stage reusable once these MMAs retire
        }
        umma_commit(accum_bar);
    }
    __syncwarp();

    // ---- epilogue: TMEM -> registers -> global ----
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + warp * 32 + lane;  // TMEM lane == accumulator row
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        const size_t off = (size_t)row * ldd + n0 + c * 32;
        if (EPI == EPI_F32) {
            float4 *dst = (float4 *)(d_f32 + off);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        } else {
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(kTmemCols) : "memory");
}

// ---- operand preparation (HBM-bound elementwise kernels) -----------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Twiddle operands for both passes; np = padded size, rows/cols beyond n are zero.
//   a1 (pass 1, K interleaved): a1[2l+c][2n+c'],   a2 (pass 2, K planar): a2[2k+c][c' np + m]
// W[k, m] = exp(sign 2 pi i ((k - so)(m - si) mod n) / n): the shifts express the fftshift /
// ifftshift pairs of the reference (see dft_kind_params).
__global__ void twiddle_kernel(int n, int np, int so, int si, int sign, __nv_bfloat16 *a1_hi,
                               __nv_bfloat16 *a1_lo, __nv_bfloat16 *a2_hi, __nv_bfloat16 *a2_lo) {
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
            const size_t i1 = (size_t)(2 * row + c) * ld + 2 * col + cp;
            const size_t i2 = (size_t)(2 * row + c) * ld + (size_t)cp * np + col;
            a1_hi[i1] = h; a1_lo[i1] = l;
            a2_hi[i2] = h; a2_lo[i2] = l;
        }
}

// data operand of pass 1 from interleaved complex64: b1[m][2n+c'] (zero padded to np)
__global__ void prep_b_c64_kernel(const float2 *__restrict__ b, int n, int np,
                                  __nv_bfloat16 *b_hi, __nv_bfloat16 *b_lo) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= np) return;
    float2 v = make_float2(0.f, 0.f);
    if (row < n && col < n) v = b[(size_t)row * n + col];
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v.x, h0, l0);
    split_bf16(v.y, h1, l1);
    const size_t i = (size_t)row * 2 * np + 2 * col;
    *(__nv_bfloat162 *)(b_hi + i) = __nv_bfloat162(h0, h1);
    *(__nv_bfloat162 *)(b_lo + i) = __nv_bfloat162(l0, l1);
}

// a2b input: |A| exp(i phi pi/180), opt_analyze.py:224 (amplitude and degrees, float64)
__global__ void prep_b_a2b_kernel(const double *__restrict__ amp, const double *__restrict__ phi,
                                  int n, int np, __nv_bfloat16 *b_hi, __nv_bfloat16 *b_lo) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= np) return;
    float2 v = make_float2(0.f, 0.f);
    if (row < n && col < n) {
        const size_t i = (size_t)row * n + col;
        double s, c;
        sincos(phi[i] * 3.141592653589793 / 180.0, &s, &c);
        const double a = fabs(amp[i]);
        v = make_float2((float)(a * c), (float)(a * s));
    }
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v.x, h0, l0);
    split_bf16(v.y, h1, l1);
    const size_t i = (size_t)row * 2 * np + 2 * col;
    *(__nv_bfloat162 *)(b_hi + i) = __nv_bfloat162(h0, h1);
    *(__nv_bfloat162 *)(b_lo + i) = __nv_bfloat162(l0, l1);
}

// d2 planar fp32 [2 np][np] (row 2k+c, column l) -> interleaved complex64 [n][n]
__global__ void finish_c64_kernel(const float *__restrict__ d2, int n, int np, float scale,
                                  float2 *__restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= n) return;
    out[(size_t)row * n + col] = make_float2(scale * d2[(size_t)(2 * row) * np + col],
                                             scale * d2[(size_t)(2 * row + 1) * np + col]);
}

// a2b outputs: complex128 beam and phase = arctan2(imag, real), opt_analyze.py:228-229
__global__ void finish_a2b_kernel(const float *__restrict__ d2, int n, int np, double scale,
                                  double2 *__restrict__ beam, double *__restrict__ phase) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= n) return;
    const double re = scale * d2[(size_t)(2 * row) * np + col];
    const double im = scale * d2[(size_t)(2 * row + 1) * np + col];
    beam[(size_t)row * n + col] = make_double2(re, im);
    phase[(size_t)row * n + col] = atan2(im, re);
}

// ---- host side -------------------------------------------------------------------------------
struct DftWorkspace {
    uint8_t *base;
    int n, np;
    __nv_bfloat16 *a1_hi, *a1_lo, *a2_hi, *a2_lo, *b1_hi, *b1_lo, *d1_hi, *d1_lo;
    float *d2;
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
    w.d2 = (float *)p;
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

// 2-D bf16 row-major [rows][cols] with 128B-swizzled {BK x 128} boxes
static int make_map(CUtensorMap *m, const void *ptr, uint64_t rows, uint64_t cols) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(SOSAT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {BK, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(SOSAT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// Twiddle operands depend only on n; they are generated once per (device, workspace, n) and
// remembered in a small host-side registry so that repeat transforms launch no extra work.
// A caller that frees a planned workspace must call sosat_dft2_forget(ptr) before the address
// can be reused for another workspace.
struct PlanEntry {
    int device;
    const void *ws;
    int n, kind;
};
static PlanEntry g_plans[64];
static int g_num_plans = 0;

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

static int ensure_twiddles(const DftWorkspace &w, int kind, cudaStream_t st) {
    int dev = 0, so, si, sign;
    double scale;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    if (int rc = dft_kind_params(kind, w.n, so, si, sign, scale)) return rc;
    for (int i = 0; i < g_num_plans; ++i)
        if (g_plans[i].device == dev && g_plans[i].ws == w.base && g_plans[i].n == w.n &&
            g_plans[i].kind == kind)
            return 0;
    dim3 grid((w.np + 127) / 128, w.np);
    twiddle_kernel<<<grid, 128, 0, st>>>(w.n, w.np, so, si, sign, w.a1_hi, w.a1_lo, w.a2_hi, w.a2_lo);
    SOSAT_CUDA_OK(cudaGetLastError());
    int slot = -1;
    for (int i = 0; i < g_num_plans; ++i)
        if (g_plans[i].device == dev && g_plans[i].ws == w.base) slot = i;  // same buffer, re-planned
    if (slot < 0) slot = g_num_plans < 64 ? g_num_plans++ : 0;
    g_plans[slot] = PlanEntry{dev, w.base, w.n, kind};
    return 0;
}

// the two GEMM passes; b1 must already hold the split data operand
static int run_passes(const DftWorkspace &w, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        SOSAT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_kernel<EPI_F32>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        SOSAT_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_kernel<EPI_SPLIT_BF16>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        attr_set = true;
    }
    const uint64_t np = w.np;
    CUtensorMap a1h, a1l, a2h, a2l, b1h, b1l, b2h, b2l;
    int rc = 0;
    if ((rc = make_map(&a1h, w.a1_hi, 2 * np, 2 * np))) return rc;
    if ((rc = make_map(&a1l, w.a1_lo, 2 * np, 2 * np))) return rc;
    if ((rc = make_map(&a2h, w.a2_hi, 2 * np, 2 * np))) return rc;
    if ((rc = make_map(&a2l, w.a2_lo, 2 * np, 2 * np))) return rc;
    if ((rc = make_map(&b1h, w.b1_hi, np, 2 * np))) return rc;
    if ((rc = make_map(&b1l, w.b1_lo, np, 2 * np))) return rc;
    // pass-2 data operand = pass-1 output [2 np][np] viewed as [np][2 np]
    if ((rc = make_map(&b2h, w.d1_hi, np, 2 * np))) return rc;
    if ((rc = make_map(&b2l, w.d1_lo, np, 2 * np))) return rc;
    const dim3 grid(w.np / BN, 2 * w.np / BM);
    const int num_kb = 2 * w.np / BK;
    gemm_bf16x3_kernel<EPI_SPLIT_BF16><<<grid, 128, kGemmSmem, st>>>(
        a1h, a1l, b1h, b1l, num_kb, nullptr, w.d1_hi, w.d1_lo, w.np);
    gemm_bf16x3_kernel<EPI_F32><<<grid, 128, kGemmSmem, st>>>(a2h, a2l, b2h, b2l, num_kb, w.d2,
                                                              nullptr, nullptr, w.np);
    SOSAT_CUDA_OK(cudaGetLastError());
    return 0;
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
    for (int i = 0; i < g_num_plans; ++i)
        if (g_plans[i].ws == d_workspace) g_plans[i] = g_plans[--g_num_plans], --i;
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
    if (int rc = ensure_twiddles(w, kind, st)) return rc;
    dim3 grid((w.np + 127) / 128, w.np);
    prep_b_c64_kernel<<<grid, 128, 0, st>>>((const float2 *)d_b, n, w.np, w.b1_hi, w.b1_lo);
    if (int rc = run_passes(w, st)) return rc;
    dim3 g2((n + 127) / 128, n);
    finish_c64_kernel<<<g2, 128, 0, st>>>(w.d2, n, w.np, (float)scale, (float2 *)d_B);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
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
    if (int rc = ensure_twiddles(w, kind, st)) return rc;
    dim3 grid((w.np + 127) / 128, w.np);
    prep_b_a2b_kernel<<<grid, 128, 0, st>>>(d_amp, d_phi_deg, n, w.np, w.b1_hi, w.b1_lo);
    if (int rc = run_passes(w, st)) return rc;
    dim3 g2((n + 127) / 128, n);
    finish_a2b_kernel<<<g2, 128, 0, st>>>(w.d2, n, w.np, scale, (double2 *)d_out, d_phase);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
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
