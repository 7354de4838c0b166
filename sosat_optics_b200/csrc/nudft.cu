// Stage 2 on irregular samples: direct-sum far field (north_star)
//     B_j = sum_p b_p exp(-2 pi i (x_p thx_j + y_p thy_j) / lambda)
// One output angle per thread (two per thread for ILP), sample points staged through shared
// memory as float4 {x, y, Re b, Im b}, phase factor generated in registers: phase in cycles is
// reduced with a round-to-nearest subtraction and fed to the SFU sin/cos (MUFU), whose
// absolute error (~5e-7) is far inside the 1e-4-of-peak tolerance.  Bound by the SFU pipe:
// 2 MUFU + ~10 FP32 ops per (point, angle) pair; no HBM traffic beyond one pass over inputs.
#include "common.cuh"

namespace sosat {

constexpr int kNuThreads = 256;
constexpr int kNuPts = 1024;   // points per shared-memory chunk (16 KiB)
constexpr int kNuPerThread = 2;

__global__ void __launch_bounds__(kNuThreads)
nudft_kernel(const float4 *__restrict__ xyb, long long n_pts, const float2 *__restrict__ th,
             long long n_ang, float inv_lambda, float2 *__restrict__ out) {
    __shared__ float4 s_pts[kNuPts];
    float tx[kNuPerThread], ty[kNuPerThread], ar[kNuPerThread], ai[kNuPerThread];
    const long long j0 = ((long long)blockIdx.x * kNuThreads + threadIdx.x);
    const long long stride = (long long)gridDim.x * kNuThreads;
#pragma unroll
    for (int q = 0; q < kNuPerThread; ++q) {
        const long long j = j0 + q * stride;
        const float2 t = j < n_ang ? th[j] : make_float2(0.f, 0.f);
        tx[q] = -t.x * inv_lambda;  // cycles per unit x, sign of the numpy FFT
        ty[q] = -t.y * inv_lambda;
        ar[q] = 0.f;
        ai[q] = 0.f;
    }
    for (long long p0 = 0; p0 < n_pts; p0 += kNuPts) {
        const int m = (int)min((long long)kNuPts, n_pts - p0);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += kNuThreads) s_pts[i] = xyb[p0 + i];
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const float4 pt = s_pts[i];
#pragma unroll
            for (int q = 0; q < kNuPerThread; ++q) {
                float cyc = fmaf(pt.x, tx[q], pt.y * ty[q]);
                // cyc -= rint(cyc) with the 1.5 * 2^23 trick: two FADDs on the FMA pipe instead
                // of an FRND on the XU pipe that the two MUFUs already saturate (|cyc| < 2^22)
                cyc -= __fadd_rn(__fadd_rn(cyc, 12582912.f), -12582912.f);
                const float ang = cyc * 6.2831853071795865f;
                const float sn = __sinf(ang), cs = __cosf(ang);
                ar[q] = fmaf(pt.z, cs, fmaf(-pt.w, sn, ar[q]));
                ai[q] = fmaf(pt.z, sn, fmaf(pt.w, cs, ai[q]));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kNuPerThread; ++q) {
        const long long j = j0 + q * stride;
        if (j < n_ang) out[j] = make_float2(ar[q], ai[q]);
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
    nudft_kernel<<<(unsigned)blocks, kNuThreads, 0, (cudaStream_t)stream>>>(
        (const float4 *)d_xyb, n_pts, (const float2 *)d_th, n_ang, (float)inv_lambda,
        (float2 *)d_B);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}
