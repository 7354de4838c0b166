// Stage 2 on irregular samples: direct-sum far field (north_star)
//     B_j = sum_p b_p exp(-2 pi i (x_p thx_j + y_p thy_j) / lambda)
// One output angle per thread (two per thread for ILP), sample points staged through shared
// memory as float4 {x, y, Re b, Im b}, phase factor generated in registers by the SFU sin/cos
// (MUFU), whose absolute error (~5e-7) is far inside the 1e-4-of-peak tolerance.  The phase is
// formed in radians (the angle tables are pre-scaled by -2 pi / lambda) and handed to
// __sinf / __cosf as it is: their FMUL.RZ by 1/2pi turns it into revolutions and MUFU.SIN/COS
// take the fractional part themselves, so an explicit range reduction (3 FADDs + 2 FMULs per
// pair in the first version) adds nothing -- the fp32 rounding of the phase itself,
// 6e-8 |phase|, is the error either way.  Bound by the SFU pipe: 2 MUFU + 7 FP32 per (point,
// angle) pair; no HBM traffic beyond one pass over the inputs.
// The grid is (angle blocks) x (point chunks): when the angles alone do not fill the machine
// (a 256^2 map is 128 blocks of 512 angles) the points are split too and the chunks' partial
// sums are combined with red.global.add.f32 into the zeroed output.
#include "common.cuh"

namespace sosat {

constexpr int kNuThreads = 256;
constexpr int kNuPts = 1024;   // points per shared-memory chunk (16 KiB)
constexpr int kNuPerThread = 2;

template <bool ATOMIC>
__global__ void __launch_bounds__(kNuThreads)
nudft_kernel(const float4 *__restrict__ xyb, long long n_pts, const float2 *__restrict__ th,
             long long n_ang, float scale, long long pts_per_split, float2 *__restrict__ out) {
    __shared__ float4 s_pts[kNuPts];
    // Two-level accumulation: fp32 within a shared-memory chunk (1024 points: coherent sums stay
    // below 2^10 |b|, so every sample keeps >= 13 bits), chunk sums folded into fp64 -- a coherent
    // beam peak over 1e7+ points per split would otherwise reach 2^23..2^24 |b| in fp32, where the
    // ulp is the sample amplitude.  4 DADD + 4 F2F per thread and chunk: 0.2 % of the chunk's work.
    float tx[kNuPerThread], ty[kNuPerThread];
    double ar[kNuPerThread], ai[kNuPerThread];
    const long long j0 = ((long long)blockIdx.x * kNuThreads + threadIdx.x);
    const long long stride = (long long)gridDim.x * kNuThreads;
#pragma unroll
    for (int q = 0; q < kNuPerThread; ++q) {
        const long long j = j0 + q * stride;
        const float2 t = j < n_ang ? th[j] : make_float2(0.f, 0.f);
        tx[q] = t.x * scale;  // radians per unit x, sign of the numpy FFT (scale = -2 pi / lambda)
        ty[q] = t.y * scale;
        ar[q] = 0.0;
        ai[q] = 0.0;
    }
    const long long p_begin = (long long)blockIdx.y * pts_per_split;
    const long long p_end = min(n_pts, p_begin + pts_per_split);
    for (long long p0 = p_begin; p0 < p_end; p0 += kNuPts) {
        const int m = (int)min((long long)kNuPts, p_end - p0);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += kNuThreads) s_pts[i] = xyb[p0 + i];
        __syncthreads();
        float cr[kNuPerThread], ci[kNuPerThread];
#pragma unroll
        for (int q = 0; q < kNuPerThread; ++q) cr[q] = ci[q] = 0.f;
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const float4 pt = s_pts[i];
#pragma unroll
            for (int q = 0; q < kNuPerThread; ++q) {
                const float ang = fmaf(pt.x, tx[q], pt.y * ty[q]);
                const float sn = __sinf(ang), cs = __cosf(ang);
                cr[q] = fmaf(pt.z, cs, fmaf(-pt.w, sn, cr[q]));
                ci[q] = fmaf(pt.z, sn, fmaf(pt.w, cs, ci[q]));
            }
        }
#pragma unroll
        for (int q = 0; q < kNuPerThread; ++q) {
            ar[q] += (double)cr[q];
            ai[q] += (double)ci[q];
        }
    }
#pragma unroll
    for (int q = 0; q < kNuPerThread; ++q) {
        const long long j = j0 + q * stride;
        if (j < n_ang) {
            if (ATOMIC) {
                atomicAdd(&out[j].x, (float)ar[q]);
                atomicAdd(&out[j].y, (float)ai[q]);
            } else {
                out[j] = make_float2((float)ar[q], (float)ai[q]);
            }
        }
    }
}

}  // namespace sosat

using namespace sosat;

extern "C" int sosat_nudft_direct(const float *d_xyb, int64_t n_pts, const float *d_th,
                                  int64_t n_ang, double inv_lambda, float *d_B, void *stream) {
    SOSAT_REQUIRE(d_xyb && d_th && d_B, "sosat_nudft_direct: null pointer");
    SOSAT_REQUIRE(n_pts >= 0 && n_ang >= 0, "negative size");
    SOSAT_REQUIRE(((uintptr_t)d_xyb & 15) == 0 && ((uintptr_t)d_th & 7) == 0 &&
                      ((uintptr_t)d_B & 7) == 0,
                  "sosat_nudft_direct: misaligned pointer");
    if (n_ang == 0) return SOSAT_OK;
    long long blocks = (n_ang + (long long)kNuThreads * kNuPerThread - 1) /
                       ((long long)kNuThreads * kNuPerThread);
    SOSAT_REQUIRE(blocks < 2147483647LL, "too many output angles for one launch");
    // split the points until the grid holds ~4 waves of 4 resident CTAs per SM (whole
    // shared-memory chunks per split; at most 64 splits)
    const long long want = (long long)num_sms() * 16;
    long long chunks = (n_pts + kNuPts - 1) / kNuPts;
    long long splits = blocks >= want ? 1 : (want + blocks - 1) / blocks;
    if (splits > chunks) splits = chunks;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    const long long pts_per_split = (chunks + splits - 1) / splits * kNuPts;
    splits = n_pts > 0 ? (n_pts + pts_per_split - 1) / pts_per_split : 1;
    const float scale = (float)(-6.283185307179586476925 * inv_lambda);
    const dim3 grid((unsigned)blocks, (unsigned)splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (splits > 1) {
        SOSAT_CUDA_OK(cudaMemsetAsync(d_B, 0, (size_t)n_ang * sizeof(float2), st));
        nudft_kernel<true><<<grid, kNuThreads, 0, st>>>((const float4 *)d_xyb, n_pts,
                                                        (const float2 *)d_th, n_ang, scale,
                                                        pts_per_split, (float2 *)d_B);
    } else {
        nudft_kernel<false><<<grid, kNuThreads, 0, st>>>((const float4 *)d_xyb, n_pts,
                                                         (const float2 *)d_th, n_ang, scale,
                                                         pts_per_split > 0 ? pts_per_split : kNuPts,
                                                         (float2 *)d_B);
    }
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}
