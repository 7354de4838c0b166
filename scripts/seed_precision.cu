// Measures the relative error of the fp64 MUFU seeds (rcp.approx.ftz.f64, rsqrt.approx.ftz.f64)
// over a range of inputs.  nvcc -arch=sm_100a scripts/seed_precision.cu -o /tmp/seedp && /tmp/seedp
#include <cstdio>
#include <cmath>
__global__ void k(const double* x, double* rc, double* rs, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a, b;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(a) : "d"(x[i]));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(b) : "d"(x[i]));
    rc[i] = a; rs[i] = b;
}
int main() {
    const int n = 1 << 20;
    double *x, *rc, *rs;
    cudaMallocManaged(&x, n * 8); cudaMallocManaged(&rc, n * 8); cudaMallocManaged(&rs, n * 8);
    const double ranges[][2] = {{0.5, 1.0}, {0.9, 1.0}, {1.0, 2.0}, {2.0, 4.0}, {0.05, 0.5}, {-1.0, -0.5}};
    for (auto& r : ranges) {
        for (int i = 0; i < n; ++i) x[i] = r[0] + (r[1] - r[0]) * (i + 0.37) / n;
        k<<<n / 256, 256>>>(x, rc, rs, n);
        cudaDeviceSynchronize();
        double e1 = 0, e2 = 0;
        for (int i = 0; i < n; ++i) {
            e1 = fmax(e1, fabs(rc[i] * x[i] - 1.0));
            if (x[i] > 0) e2 = fmax(e2, fabs(rs[i] * sqrt(x[i]) - 1.0));
        }
        printf("x in [%g,%g]: max rel err rcp seed %.3e  rsqrt seed %.3e\n", r[0], r[1], e1, e2);
    }
    return 0;
}
