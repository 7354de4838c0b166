// B200 instruction latency / throughput probes used for the K1 cost model.
//   dependent-chain latency of DFMA, FFMA; throughput of MUFU.RSQ64H, MUFU.RCP64H, F2F f64<->f32
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_latency scripts/dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x + 1.0, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    float y0 = threadIdx.x + 1.0f, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
    const float fa = (float)a, fb = (float)b;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) x0 = fma(x0, a, b);                       // DFMA dependent chain
            if (MODE == 1) y0 = fmaf(y0, fa, fb);                    // FFMA dependent chain
            if (MODE == 2) {                                         // 2 independent DFMA chains
                x0 = fma(x0, a, b); x1 = fma(x1, a, b);
            }
            if (MODE == 3) {                                         // MUFU.RSQ64H throughput
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x0));
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x1));
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x2));
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x3));
            }
            if (MODE == 4) {                                         // MUFU.RCP64H throughput
                asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x0));
                asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x1));
                asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x2));
                asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x3));
            }
            if (MODE == 5) {                                         // F2F.F32.F64 + F2F.F64.F32 round trips
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y0) : "d"(x0));
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x0) : "f"(y0));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y1) : "d"(x1));
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x1) : "f"(y1));
            }
            if (MODE == 6) {                                         // F2F.F32.F64 only (4 independent)
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y0) : "d"(x0));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y1) : "d"(x1));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y2) : "d"(x2));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y3) : "d"(x3));
                x0 += 1.0; x1 += 1.0; x2 += 1.0; x3 += 1.0;  // 4 DADD to keep the inputs changing
            }
            if (MODE == 7) {                                         // DADD only (baseline of mode 6)
                x0 += 1.0; x1 += 1.0; x2 += 1.0; x3 += 1.0;
            }
            if (MODE == 8) {                                         // DMUL dependent chain
                x0 = x0 * a;
            }
            if (MODE == 9) {                                         // MUFU.RSQ dependent chain latency
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y0));
            }
            if (MODE == 10) {                                        // RSQ64H dependent chain latency
                asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x0));
            }
        }
    }
    double s = x0 + x1 + x2 + x3 + (double)(y0 + y1 + y2 + y3);
    if (s == 123.456) out[0] = s;
}

template <int MODE>
double run(double *d, int iters, int blocks, int threads) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    return best * 1e-3 * 1.965e9;  // cycles
}

int main() {
    double *d; cudaMalloc(&d, 64);
    const int iters = 1 << 12;
    const double n = (double)iters * 16;
    // latency: one warp per SMSP (148 blocks x 128 threads)
    printf("latency (1 warp/SMSP, dependent chain), cycles per instruction:\n");
    printf("  DFMA %.2f  DMUL %.2f  FFMA %.2f  2xDFMA(indep) %.2f per pair  MUFU.RSQ %.2f  MUFU.RSQ64H %.2f\n",
           run<0>(d, iters, 148, 128) / n, run<8>(d, iters, 148, 128) / n, run<1>(d, iters, 148, 128) / n,
           run<2>(d, iters, 148, 128) / n, run<9>(d, iters, 148, 128) / n, run<10>(d, iters, 148, 128) / n);
    // throughput: 8 warps per SMSP (148 x 4 blocks x 256 threads)
    const double w = 8;
    printf("throughput (8 warps/SMSP), SMSP cycles per warp instruction:\n");
    printf("  MUFU.RSQ64H %.2f  MUFU.RCP64H %.2f  F2F pair(f64->f32->f64) %.2f per pair\n",
           run<3>(d, iters, 592, 256) / (n * w * 4), run<4>(d, iters, 592, 256) / (n * w * 4),
           run<5>(d, iters, 592, 256) / (n * w * 2));
    const double c6 = run<6>(d, iters, 592, 256) / (n * w * 4), c7 = run<7>(d, iters, 592, 256) / (n * w * 4);
    printf("  F2F.F32.F64 + DADD %.2f per pair; DADD alone %.2f\n", c6, c7);
    return 0;
}
