// Issue cost of the packed fp32 FMA (FFMA2, sm_100): does one FFMA2 take one issue slot, and
// does it still do so next to DFMAs of the same warp?  Modes: 0 FFMA x8, 1 FFMA2 x4 (same flops),
// 2 DFMA x4 + FFMA x8, 3 DFMA x4 + FFMA2 x4.   nvcc -arch=sm_100a -O3 -o ffma2_issue scripts/ffma2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, float fa, float fb) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    float2 y0 = make_float2(threadIdx.x, threadIdx.x + 1.f), y1 = make_float2(y0.x + 2.f, y0.x + 3.f),
           y2 = make_float2(y0.x + 4.f, y0.x + 5.f), y3 = make_float2(y0.x + 6.f, y0.x + 7.f);
    const float2 A = make_float2(fa, fa * 1.0001f), B = make_float2(fb, fb * 1.0001f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // modes 4 / 5: even warps run only the DFMAs, odd warps only the FFMA / FFMA2 stream
            const bool dp_warp = ((threadIdx.x >> 7) & 1) == 0;  // warps 0-3 | 4-7: one of each kind per SMSP
            if ((MODE == 2 || MODE == 3) || (MODE >= 4 && dp_warp)) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            }
            if (MODE >= 4 && dp_warp) continue;
            if (MODE == 0 || MODE == 2 || MODE == 4) {
                y0.x = fmaf(y0.x, A.x, B.x); y0.y = fmaf(y0.y, A.y, B.y);
                y1.x = fmaf(y1.x, A.x, B.x); y1.y = fmaf(y1.y, A.y, B.y);
                y2.x = fmaf(y2.x, A.x, B.x); y2.y = fmaf(y2.y, A.y, B.y);
                y3.x = fmaf(y3.x, A.x, B.x); y3.y = fmaf(y3.y, A.y, B.y);
            } else {
                y0 = __ffma2_rn(y0, A, B); y1 = __ffma2_rn(y1, A, B);
                y2 = __ffma2_rn(y2, A, B); y3 = __ffma2_rn(y3, A, B);
            }
        }
    }
    double s = x0 + x1 + x2 + x3 + (double)(y0.x + y1.x + y2.x + y3.x + y0.y + y1.y + y2.y + y3.y);
    if (s == 123.456) out[0] = s;
}

template <int MODE>
float run(double *d, int iters, int warps_per_smsp) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * warps_per_smsp / 2;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, iters, 0.999999, 1e-9, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    return best;
}

int main() {
    double *d; cudaMalloc(&d, 64);
    const int iters = 1 << 13;
    for (int w : {4, 8, 16}) {
        float t0 = run<0>(d, iters, w), t1 = run<1>(d, iters, w), t2 = run<2>(d, iters, w), t3 = run<3>(d, iters, w);
        float t4 = run<4>(d, iters, w), t5 = run<5>(d, iters, w);
        double n = (double)w * iters * 8;  // per SMSP: groups of (8 FFMA | 4 FFMA2 [+ 4 DFMA])
        const double c = 1e-3 * 1.965e9 / n;
        printf("warps/SMSP %2d: cycles per group: 8 FFMA %.2f | 4 FFMA2 %.2f | 4 DFMA + 8 FFMA %.2f | 4 DFMA + 4 FFMA2 %.2f\n",
               w, t0 * c, t1 * c, t2 * c, t3 * c);
        // split modes: w/2 warps of each kind per SMSP -> per pair of warps one DP group and one fp32 group
        printf("               split warps (per DP-group + fp32-group of two different warps): DFMA | FFMA %.2f   DFMA | FFMA2 %.2f\n",
               t4 * c * 2, t5 * c * 2);
    }
    return 0;
}
