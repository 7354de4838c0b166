// On-device reductions of the near and far field, so that frequency
// and receiver sweeps return a few scalars per problem instead of whole fields:
//   sosat_near_field_metrics  get_beam_cent, get_fwhm, get_angle_out's tan_og   ray_trace.py:1652-1672
//   sosat_ff_fwhm_indices     the peak / half-power search of ff_fwhm_est        ray_trace.py:1757-1779
// All of it is O(rays) or O(N^2) streaming work over arrays that already sit in HBM.
#include <cstring>

#include "common.cuh"

namespace sosat {

static constexpr double kHalfPi = 1.5707963267948966;
static constexpr double kEightLn2 = 5.545177444479562;  // a^2 = exp(-8 ln2 ((de_ho/fx)^2 + (de_ve/fy)^2))

struct MetricFreq {
    double c2x[SOSAT_MAX_FREQS];  // -8 ln2 / fwhp_x^2
    double c2y[SOSAT_MAX_FREQS];
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// non-negative doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(double *dst, double v) {
    atomicMax((unsigned long long *)dst, (unsigned long long)__double_as_longlong(v));
}

// numpy.linspace(start, stop, n)[i], as in trace.cu
__device__ __forceinline__ double lin_at(double start, double stop, double step, int n, int i) {
    if (i == n - 1 && n > 1) return stop;
    return __dadd_rn(__dmul_rn((double)i, step), start);
}

// work: [0] n, [1] sum x, [2] sum z, [3..5] sum exit direction, [8 + f] max a_f^2
__global__ void nfm_sums_kernel(const double *__restrict__ out, long long n, double *work) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        if (out[4 * n + i] != 0.0) {  // xx = np.where(out[4] != 0), ray_trace.py:1599
            acc[0] += 1.0;
            acc[1] += out[0 * n + i];
            acc[2] += out[2 * n + i];
            acc[3] += out[8 * n + i];
            acc[4] += out[9 * n + i];
            acc[5] += out[10 * n + i];
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double s = warp_sum_d(acc[q]);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(work + q, s);
    }
}

struct NfmParams {
    double lin_start, lin_stop, lin_step;
    long long ray_begin, n;
    int n_scan, n_freq;
};

// squared launch-angle offsets of ray ii: de_ho = pi/2 - theta, de_ve = phi - pi/2
__device__ __forceinline__ void launch_offsets(const NfmParams &p, long long i, double &h2, double &v2) {
    const long long ii = p.ray_begin + i;
    const int ith = (int)(ii % p.n_scan), iph = (int)(ii / p.n_scan);
    const double ho = lin_at(p.lin_start, p.lin_stop, p.lin_step, p.n_scan, ith) - kHalfPi;
    const double ve = lin_at(p.lin_start, p.lin_stop, p.lin_step, p.n_scan, iph) - kHalfPi;
    h2 = ho * ho;
    v2 = ve * ve;
}

__global__ void nfm_amax_kernel(const double *__restrict__ out, NfmParams p, MetricFreq mf,
                                double *work) {
    for (int f = 0; f < p.n_freq; ++f) {
        double best = 0.0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n;
             i += (long long)gridDim.x * blockDim.x) {
            if (out[4 * p.n + i] != 0.0) {
                double h2, v2;
                launch_offsets(p, i, h2, v2);
                best = fmax(best, exp(fma(mf.c2x[f], h2, mf.c2y[f] * v2)));
            }
        }
        best = warp_max_d(best);
        if ((threadIdx.x & 31) == 0 && best > 0.0) atomic_max_nonneg(work + 8 + f, best);
    }
}

// result: [0] n_valid, [1] mean x_sim [cm], [2] mean y_sim [cm], [3..5] tan_og,
//         [6 + 2f] fwhm_nf_x = max |y_sim - cy|, [7 + 2f] fwhm_nf_y = max |x_sim - cx| over rays
//         with a_f^2 > max(a_f^2)/2  (the reference's swapped naming, ray_trace.py:1660-1665)
__global__ void nfm_fwhm_kernel(const double *__restrict__ out, NfmParams p, MetricFreq mf,
                                const double *__restrict__ work, double *result) {
    const double nv = work[0];
    const double cx = nv > 0 ? work[1] / nv / 1e1 : 0.0, cy = nv > 0 ? work[2] / nv / 1e1 : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        result[0] = nv;
        result[1] = cx;
        result[2] = cy;
        for (int q = 0; q < 3; ++q) result[3 + q] = nv > 0 ? work[3 + q] / nv : 0.0;
    }
    for (int f = 0; f < p.n_freq; ++f) {
        const double half = work[8 + f] / 2;
        double mx = 0.0, my = 0.0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n;
             i += (long long)gridDim.x * blockDim.x) {
            if (out[4 * p.n + i] != 0.0) {
                double h2, v2;
                launch_offsets(p, i, h2, v2);
                if (exp(fma(mf.c2x[f], h2, mf.c2y[f] * v2)) > half) {
                    my = fmax(my, fabs(out[2 * p.n + i] / 1e1 - cy));
                    mx = fmax(mx, fabs(out[0 * p.n + i] / 1e1 - cx));
                }
            }
        }
        my = warp_max_d(my);
        mx = warp_max_d(mx);
        if ((threadIdx.x & 31) == 0) {
            if (my > 0.0) atomic_max_nonneg(result + 6 + 2 * f, my);
            if (mx > 0.0) atomic_max_nonneg(result + 7 + 2 * f, mx);
        }
    }
}

// ---- far field: peak of |B|^2 and the half-power extent through the peak row and column ------
// One launch covers a whole stack of fields (blockIdx.y = field): a receiver x frequency sweep
// reads back ONE small array instead of synchronising once per field.
// packed = float bits of |B|^2 << 32 | (0xffffffff - flat index): atomicMax picks the largest
// value and, among equal values, the smallest index (np.where(...)[0][0], row-major).
__global__ void ff_init_kernel(int n, int n_fields, unsigned long long *packed, int *idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_fields) {
        packed[i] = 0ull;
        int *d = idx + 6 * i;
        d[0] = 0; d[1] = 0; d[2] = n; d[3] = -1; d[4] = n; d[5] = -1;
    }
}

__global__ void ff_peak_kernel(const float2 *__restrict__ B_all, long long n2,
                               unsigned long long *packed) {
    const float2 *B = B_all + (long long)blockIdx.y * n2;
    unsigned long long best = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
         i += (long long)gridDim.x * blockDim.x) {
        const float2 v = B[i];
        const float a = v.x * v.x + v.y * v.y;
        if (a == a) {
            const unsigned long long pk = ((unsigned long long)__float_as_uint(a) << 32) |
                                          (unsigned long long)(0xffffffffu - (unsigned)i);
            best = pk > best ? pk : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(packed + blockIdx.y, best);
}

// idx (per field): [0] peak row, [1] peak column, [2],[3] first/last column above half power in
// the peak row, [4],[5] first/last row above half power in the peak column; val = peak |B|^2.
// A field without a finite non-zero cell (one NaN in b spreads through the contraction to every
// cell of B) reports all six indices as -1 and val = 0 -- the wrappers raise -- instead of
// decoding the untouched 0 into a row index far outside the array.
__global__ void ff_extent_kernel(const float2 *__restrict__ B_all, int n,
                                 const unsigned long long *__restrict__ packed, int *idx_all,
                                 float *val) {
    const float2 *B = B_all + (size_t)blockIdx.y * n * n;
    int *idx = idx_all + 6 * blockIdx.y;
    const unsigned long long pk = packed[blockIdx.y];
    const unsigned flat = 0xffffffffu - (unsigned)(pk & 0xffffffffull);
    const float peak = __uint_as_float((unsigned)(pk >> 32));
    const int pr = (int)(flat / (unsigned)n), pc = (int)(flat % (unsigned)n);
    const bool along_row = blockIdx.x == 0;  // block 0: the peak row, block 1: the peak column
    if (pk == 0ull || pr >= n || !(peak > 0.f)) {
        if (threadIdx.x < 3) idx[3 * blockIdx.x + threadIdx.x] = -1;
        if (threadIdx.x == 0 && along_row) val[blockIdx.y] = 0.f;
        return;
    }
    if (threadIdx.x == 0 && along_row) {
        idx[0] = pr;
        idx[1] = pc;
        val[blockIdx.y] = peak;
    }
    int first = n, last = -1;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float2 v = along_row ? B[(size_t)pr * n + j] : B[(size_t)j * n + pc];
        // y = abs(a)/max(abs(a)) > 0.5 with a = |B|^2 / max|B|^2 (ray_trace.py:1758-1768)
        if ((v.x * v.x + v.y * v.y) / peak > 0.5f) {
            first = min(first, j);
            last = max(last, j);
        }
    }
    atomicMin(idx + (along_row ? 2 : 4), first);
    atomicMax(idx + (along_row ? 3 : 5), last);
}

}  // namespace sosat

using namespace sosat;

extern "C" int sosat_near_field_metrics(const double *d_out, int64_t n, int n_scan,
                                        int64_t ray_begin, double cone, const sosat_freq_t *h_freq,
                                        int n_freq, double *d_result, double *d_work,
                                        void *stream) {
    SOSAT_REQUIRE(d_out && h_freq && d_result && d_work, "sosat_near_field_metrics: null pointer");
    SOSAT_REQUIRE(n >= 0 && n_scan >= 1 && ray_begin >= 0 &&
                      ray_begin + n <= (int64_t)n_scan * n_scan,
                  "ray range outside the %d x %d launch grid", n_scan, n_scan);
    SOSAT_REQUIRE(n_freq >= 1 && n_freq <= SOSAT_MAX_FREQS, "n_freq must be in [1, %d]",
                  SOSAT_MAX_FREQS);
    cudaStream_t st = (cudaStream_t)stream;
    SOSAT_CUDA_OK(cudaMemsetAsync(d_work, 0, sizeof(double) * (8 + SOSAT_MAX_FREQS), st));
    SOSAT_CUDA_OK(cudaMemsetAsync(d_result, 0, sizeof(double) * (6 + 2 * SOSAT_MAX_FREQS), st));
    if (n == 0) return SOSAT_OK;
    MetricFreq mf;
    memset(&mf, 0, sizeof mf);
    for (int f = 0; f < n_freq; ++f) {
        SOSAT_REQUIRE(h_freq[f].fwhp_x != 0 && h_freq[f].fwhp_y != 0, "feed FWHP must be non-zero");
        mf.c2x[f] = -kEightLn2 / (h_freq[f].fwhp_x * h_freq[f].fwhp_x);
        mf.c2y[f] = -kEightLn2 / (h_freq[f].fwhp_y * h_freq[f].fwhp_y);
    }
    NfmParams p;
    p.lin_start = kHalfPi - cone;
    p.lin_stop = kHalfPi + cone;
    p.lin_step = n_scan > 1 ? (p.lin_stop - p.lin_start) / (double)(n_scan - 1) : 0.0;
    p.ray_begin = ray_begin;
    p.n = n;
    p.n_scan = n_scan;
    p.n_freq = n_freq;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    nfm_sums_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_out, n, d_work);
    nfm_amax_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_out, p, mf, d_work);
    nfm_fwhm_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_out, p, mf, d_work, d_result);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_ff_fwhm_indices_batch(const float *d_B, int n, int n_fields, int *d_idx,
                                           float *d_val, void *d_work, void *stream) {
    SOSAT_REQUIRE(d_B && d_idx && d_val && d_work, "sosat_ff_fwhm_indices: null pointer");
    SOSAT_REQUIRE(n >= 1 && n <= 65535, "n=%d out of range [1, 65535]", n);
    SOSAT_REQUIRE(n_fields >= 0 && n_fields <= 65535, "n_fields=%d out of range [0, 65535]", n_fields);
    if (n_fields == 0) return SOSAT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ff_init_kernel<<<(n_fields + 255) / 256, 256, 0, st>>>(n, n_fields, (unsigned long long *)d_work,
                                                          d_idx);
    const long long n2 = (long long)n * n;
    long long blocks = (n2 + 255) / 256;
    long long cap = (long long)num_sms() * 8 / n_fields;
    if (cap < 8) cap = 8;
    if (blocks > cap) blocks = cap;
    ff_peak_kernel<<<dim3((unsigned)blocks, (unsigned)n_fields), 256, 0, st>>>(
        (const float2 *)d_B, n2, (unsigned long long *)d_work);
    ff_extent_kernel<<<dim3(2, (unsigned)n_fields), 256, 0, st>>>(
        (const float2 *)d_B, n, (const unsigned long long *)d_work, d_idx, d_val);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_ff_fwhm_indices(const float *d_B, int n, int *d_idx, float *d_val,
                                     void *d_work, void *stream) {
    return sosat_ff_fwhm_indices_batch(d_B, n, 1, d_idx, d_val, d_work, stream);
}
