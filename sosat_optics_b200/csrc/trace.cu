// Stage 1 of the hot path on sm_100a: ray trace feed -> source plane (K1), in three forms:
//   trace_rows_kernel   the reference's out[17, n] table            (ray_trace.py:1070-1591)
//   trace_bin_kernel    fused trace + near-field binning b(x,y)      (north_star / SURVEY 8 a7)
//   near-field arrays   getNearField's means, phase wrap, scaling    (ray_trace.py:1594-1614)
// One ray per thread, fp64 state, geometry in __constant__ memory, no per-ray global reads.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "trace_device.cuh"

namespace sosat {

__constant__ GeomC g_geom;
static bool g_geom_set[64] = {false};
// which build of the kernels this device's geometry needs (GeoBuild below): 0 = the SAT
// prescription, 1 = any other geometry, 2 = near-parabolic surfaces (u c'/(1+s) form)
static int g_geom_build[64] = {0};

static constexpr double kHalfPi = 1.5707963267948966;
static constexpr double kFourLn2 = 2.772588722239781;  // (1/2) / (1/sqrt(8 ln 2))^2

// numpy.linspace(start, stop, n)[i] = i*step + start, last element forced to stop
// (two roundings, no fused multiply-add), ray_trace.py:1086-1087.
__device__ __forceinline__ double linspace_at(double start, double stop, double step, int n,
                                              int i) {
    if (i == n - 1 && n > 1) return stop;
    return __dadd_rn(__dmul_rn((double)i, step), start);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV per-thread doubles; one atomicAdd per value per block.
template <int NV>
__device__ __forceinline__ void block_accumulate(const double (&v)[NV], double *dst) {
    __shared__ double red[NV][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) red[i][wid] = s;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = lane < nw ? red[i][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0 && s != 0.0) atomicAdd(dst + i, s);
        }
    }
}

struct RowsParams {
    double rx[3];
    double fwhp_x, fwhp_y;
    double lin_start, lin_stop, lin_step;
    double opl_ref, kk;  // compact mode: phase = 2 pi frac(kk (OPL - opl_ref)), kk = k/2e3/(2 pi)
    int n_scan;
    long long ray_begin, n;
};

// ---- K1a: reference table out[17][n] ----------------------------------------------------
// GEO selects the surface-evaluation build (trace_device.cuh): 0 = division-free conic, only
// lens 2b refined (the SAT prescription); 1 = division-free, every surface refined (any other
// geometry); 2 = the u c'/(1+s) form (near-parabolic surfaces).
template <int GEO> struct GeoBuild {
    static constexpr bool generic = GEO == 2;
    static constexpr int mask = GEO == 0 ? 0x04 : 0x3f;  // surfaces with |c'/K'| > 300 mm
    static constexpr int edge = GEO == 0 ? 0x10 : 0x3f;  // surfaces with K' > 0
};

// COMPACT = false: the reference's table out[17][n] (fp64 rows, coalesced 8-byte row stores).
// COMPACT = true:  one float4 {x_sim [cm], y_sim [cm], a cos p, a sin p} per ray (16 B/ray, one
//                  coalesced 16-byte store; invalid rays are all-zero samples) -- the input
//                  format of the direct-sum far field (nudft.cu), SURVEY 8d "compact SoA mode".
template <int N32, int N64, int GEO, bool COMPACT>
__global__ void __launch_bounds__(256)
trace_rows_kernel(RowsParams p, double *__restrict__ out, float4 *__restrict__ samples,
                  double *__restrict__ stats) {
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live_thread = j0 < p.n;
    const long long j = live_thread ? j0 : p.n - 1;  // idle lanes shadow the last ray: whole warps trace
    double acc[SOSAT_NUM_STATS] = {0, 0, 0, 0, 0, 0, 0, 0};
    {
        const long long ii = p.ray_begin + j;
        const int ith = (int)(ii % p.n_scan), iph = (int)(ii / p.n_scan);
        const double th = linspace_at(p.lin_start, p.lin_stop, p.lin_step, p.n_scan, ith);
        const double ph = linspace_at(p.lin_start, p.lin_stop, p.lin_step, p.n_scan, iph);
        double sth, cth, sph, cph;
        sincos(th, &sth, &cth);
        sincos(ph, &sph, &cph);
        // r_hat, ray_trace.py:1103
        const double d0x = sth * cph, d0y = sth * sph, d0z = cth;
        RayResult r;
        unsigned n_slow = 0;
        trace_ray<N32, N64, GeoBuild<GEO>::generic, GeoBuild<GEO>::mask, GeoBuild<GEO>::edge>(g_geom, p.rx[0], p.rx[1], p.rx[2], d0x, d0y, d0z, r, n_slow);
        double row[11];
        if (r.flags & RAY_CLIPPED_L3) {
#pragma unroll
            for (int q = 0; q < 11; ++q) row[q] = 0.0;
        } else {
            row[0] = r.ax; row[1] = r.ay; row[2] = r.az;
            row[3] = 0.0; row[4] = 0.0;
            if (r.valid) {
                // launch angles and Gaussian feed taper, ray_trace.py:1529-1532, 1573-1577
                const double de_ve = atan(d0x / (-d0y));
                const double de_ho = atan(d0z / sqrt(d0x * d0x + d0y * d0y));
                const double hx = de_ho / p.fwhp_x, hy = de_ve / p.fwhp_y;
                row[3] = r.opl;
                row[4] = exp(-kFourLn2 * (hx * hx + hy * hy));
                const bool count = live_thread && row[4] != 0.0;
                acc[SOSAT_STAT_N_VALID] = count ? 1.0 : 0.0;
                acc[SOSAT_STAT_SUM_OPL] = count ? r.opl : 0.0;
                acc[SOSAT_STAT_SUM_DX] = count ? r.dx : 0.0;
                acc[SOSAT_STAT_SUM_DY] = count ? r.dy : 0.0;
                acc[SOSAT_STAT_SUM_DZ] = count ? r.dz : 0.0;
            }
            row[5] = r.nx; row[6] = r.ny; row[7] = r.nz;
            row[8] = r.dx; row[9] = r.dy; row[10] = r.dz;
        }
        acc[SOSAT_STAT_N_SLOW] = live_thread ? (double)n_slow : 0.0;
        acc[SOSAT_STAT_N_BRACKET_FAIL] = (live_thread && (r.flags & RAY_BRACKET_FAIL)) ? 1.0 : 0.0;
        if (live_thread && COMPACT) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row[4] != 0.0) {
                double cyc = (row[3] - p.opl_ref) * p.kk;
                cyc -= rint(cyc);  // phase reduced to [-1/2, 1/2] cycles in fp64
                float sn, cs;
                sincospif(2.0f * (float)cyc, &sn, &cs);
                const float a = (float)row[4];
                v = make_float4((float)(row[0] / 1e1), (float)(row[2] / 1e1), a * cs, a * sn);
            }
            samples[j] = v;
        }
        if (live_thread && !COMPACT) {
#pragma unroll
            for (int q = 0; q < 11; ++q) out[(long long)q * p.n + j] = row[q];
#pragma unroll
            for (int q = 11; q < SOSAT_OUT_ROWS; ++q) out[(long long)q * p.n + j] = 0.0;
        }
    }
    if (stats) block_accumulate<SOSAT_NUM_STATS>(acc, stats);
}

// ---- K1b: fused trace + binning ------------------------------------------------------------
constexpr int kTile = 16;  // rays per tile edge in theta
// A CTA of kBinWarps warps traces a 16 (theta) x kTileH (phi) patch of the launch grid, one 8 x 4
// sub-patch per warp.  Measured (1e9 rays): 8 warps 1.74e10 rays/s, 16 warps +0.3 %, 4 warps
// -1.6 %, 2 warps -4 %, barrier-free per-warp patches -4 %: warps that run in step share
// instruction-cache lines of the straight-line hot path, which outweighs the barrier skew.
#ifndef SOSAT_BIN_WARPS
#define SOSAT_BIN_WARPS 8
#endif
constexpr int kBinWarps = SOSAT_BIN_WARPS;
constexpr int kBinThreads = 32 * kBinWarps;
constexpr int kTileH = 2 * kBinWarps;
static_assert(kTileH == SOSAT_TRACE_TILE_ROWS, "sosat_trace_bin_tiles documents 16-row tiles");
// (A shared-memory bin window per CTA -- north_star's "shuffle, then shared-memory atomics" --
// was built and measured at -8 %: after the run-length aggregation a warp issues only 3-6
// global reductions per 32 rays.  Two rays per thread, skewed by half a surface so that the fp32
// steps of one ray sit next to the fp64 step of the other: -0.8 %, 128 registers.  DESIGN.md
// section 3; the code is in the history, commits 3515e61 and 95805ee.)
constexpr int kTileW = kTile;
// seed history of the predicted-seed tracer: 2 tiles x 6 surfaces x one double per thread
constexpr int kHistBytes = 2 * SOSAT_NUM_SURFACES * kBinThreads * (int)sizeof(double);

struct BinFreq {
    float neg_c_fx[SOSAT_MAX_FREQS];  // -4 ln2 / fwhp_x^2
    float neg_c_fy[SOSAT_MAX_FREQS];
    double kk[SOSAT_MAX_FREQS];       // k / 2e3 / (2 pi): OPL [mm] -> phase in cycles
};
__constant__ BinFreq g_freq;
// work queue of the persistent binning kernel: the next strip to hand out (zeroed in stream order
// before every launch; like g_freq, per-device state of one stream at a time)
__device__ unsigned long long g_next_strip;
#ifdef SOSAT_STRIP_TIMING
__device__ long long g_strip_t0[1 << 17], g_strip_t1[1 << 17];
__device__ int g_strip_sm[1 << 17];
#endif

struct BinParams {
    double lin_start, lin_stop, lin_step;
    double opl_ref, half_extent, inv_cell;
    long long row_begin;
    int n_scan, n_rows, n_rx, n_freq, grid_n;
    int tiles_x, tiles_y;
    int strip_len, strips_x;  // a strip = strip_len adjacent tiles along theta (the last may be shorter)
    int predict;              // seeds of a tile predicted from the lane's last two tiles
    long long n_strips;
    const int *tile_rows;     // optional: tile row ty of this launch is row tile_rows[ty] of the grid
};

// Persistent CTAs walk STRIPS of adjacent 16 x 16 patches of the launch grid (strips handed out
// dynamically, rim of the cone first); a warp owns an 8 (theta) x 4 (phi) sub-patch, so its 32
// rays land in a handful of neighbouring bins.
// * Seeds: from the third tile of a strip on, a warp whose 32 lanes all hold a finite two-tile
//   history takes the predicted-seed tracer (trace_device.cuh: one fp64 Newton step from
//   2 t(x-1) - t(x-2)); the first two tiles, ragged edge tiles, warps with a lane whose history
//   is NaN (outside the conic domain of lens 1b) or was re-traced, and launch grids too coarse
//   for the extrapolation (p.predict = 0) take the fp32 Newton + Halley pre-iterations.
// * Tables: sin / cos of the 16 phi rows once per strip, of the 16 theta columns per tile, double
//   buffered (the columns of tile x + 1 are filled while tile x is traced): ONE barrier per tile.
// * Binning: run-length aggregation along the lane order: every maximal run of consecutive lanes
//   with the same bin is summed into its first lane by a segmented shuffle reduction (5 steps
//   whatever the ray density), and that lane issues one red.global.add.f32 per plane.  Because
//   the optical map (theta, phi) -> (x, z) is smooth, the runs are long: 1-2 per warp at 240
//   rays/bin (configs[4]), one per row of 8 at 15 rays/bin.  Any key sequence is handled
//   correctly -- equal bins that are not adjacent in lane order simply become separate atomics.
template <int N32, int N64, int MINB, int GEO>
__global__ void __launch_bounds__(kBinThreads, MINB * 8 / kBinWarps)
trace_bin_kernel(BinParams p, const double *__restrict__ d_rx, float *__restrict__ re,
                 float *__restrict__ im, float *__restrict__ cnt, double *__restrict__ stats) {
    constexpr bool kCanPredict = N32 == 2 && N64 == 1;
    extern __shared__ double s_hist[];  // [2][surface][thread], only when p.predict
    __shared__ double s_sin_t[2][kTile], s_cos_t[2][kTile], s_sin_p[kTileH], s_cos_p[kTileH], s_rx[4];
    // Launch-angle tables, filled once per CTA: sin/cos of  l step (l < 16),  16 j step (j < 64)
    // and  start + 1024 k step (k < 64).  The angle of grid index i = l + 16 j + 1024 k is
    // start + i step (numpy.linspace, ray_trace.py:1086-1087); its sine and cosine follow from
    // two angle additions (8 DFMA/DMUL, ~3e-16 relative) instead of an fp64 sincos per tile
    // edge on the critical path in front of the tile barrier.
    __shared__ double2 s_rot[16 + 64 + 64];
    __shared__ float s_amp_t[2][SOSAT_MAX_FREQS][kTile];  // [buffer][freq][column in tile]
    __shared__ float s_amp_p[SOSAT_MAX_FREQS][kTileH];    // [freq][row in tile]
    // Running statistics.  fp64 sums: one shared-memory slot per thread (no conflicts, no
    // registers held across the tracer).  Counters: warp ballots added by lane 0 to its warp's
    // shared slot (a 64-bit shared atomicAdd is a CAS loop: 25 instructions per warp and ray,
    // measured); the rare slow-path events use plain shared atomics.
    __shared__ double s_sum[4][kBinThreads];
    __shared__ unsigned s_nvalid[kBinWarps], s_nbinned[kBinWarps];  // one slot per warp: no atomics
    __shared__ unsigned s_rare[2];                 // n_slow, n_bracket_fail
    __shared__ long long s_next;                   // the strip this CTA works on next

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lth = (lane & 7) + 8 * (warp & 1);
    const int lph = (lane >> 3) + 4 * (warp >> 1);

#pragma unroll
    for (int q = 0; q < 4; ++q) s_sum[q][tid] = 0.0;
    if (tid < kBinWarps) s_nvalid[tid] = s_nbinned[tid] = 0u;
    if (tid == 0) s_rare[0] = s_rare[1] = 0u;
    const bool use_rot = p.n_scan <= 65536;
    if (use_rot) {
        for (int e = tid; e < 16 + 64 + 64; e += kBinThreads) {
            const double a = e < 16 ? (double)e * p.lin_step
                             : e < 80 ? (double)(16 * (e - 16)) * p.lin_step
                                      : fma((double)(1024 * (e - 80)), p.lin_step, p.lin_start);
            double sn, cs;
            sincos(a, &sn, &cs);
            s_rot[e] = make_double2(sn, cs);
        }
    }
    int acc_rx = -1;
    const long long plane = (long long)p.grid_n * p.grid_n;

    auto flush_stats = [&](int irx_done) {
        __syncthreads();  // all counter updates of the finished receiver are in
        const bool t0 = tid == 0;
        unsigned nv = 0u, nb = 0u;
        if (t0) {
#pragma unroll
            for (int w = 0; w < kBinWarps; ++w) {
                nv += s_nvalid[w];
                nb += s_nbinned[w];
            }
        }
        double acc[SOSAT_NUM_STATS];
        acc[SOSAT_STAT_N_VALID] = (double)nv;
        acc[SOSAT_STAT_SUM_OPL] = s_sum[0][tid];
        acc[SOSAT_STAT_SUM_DX] = s_sum[1][tid];
        acc[SOSAT_STAT_SUM_DY] = s_sum[2][tid];
        acc[SOSAT_STAT_SUM_DZ] = s_sum[3][tid];
        acc[SOSAT_STAT_N_SLOW] = t0 ? (double)s_rare[0] : 0.0;
        acc[SOSAT_STAT_N_BRACKET_FAIL] = t0 ? (double)s_rare[1] : 0.0;
        acc[SOSAT_STAT_N_BINNED] = (double)nb;
        if (stats) block_accumulate<SOSAT_NUM_STATS>(acc, stats + (long long)irx_done * SOSAT_NUM_STATS);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) s_sum[q][tid] = 0.0;
        if (tid < kBinWarps) s_nvalid[tid] = s_nbinned[tid] = 0u;
        if (t0) s_rare[0] = s_rare[1] = 0u;
    };

    // sin / cos / Gaussian feed factors of launch-grid index idx (clamped to the grid edge: lanes
    // past it trace the edge ray so that whole warps stay converged; their result is dropped)
    auto fill_entry = [&](int idx, double &sn, double &cs, float &a2) {
        const int ic = idx < p.n_scan ? idx : p.n_scan - 1;
        const double a = linspace_at(p.lin_start, p.lin_stop, p.lin_step, p.n_scan, ic);
        if (use_rot) {
            const double2 ra = s_rot[ic & 15], rb = s_rot[16 + ((ic >> 4) & 63)],
                          rc = s_rot[80 + (ic >> 10)];
            const double s1 = fma(rc.x, rb.y, rc.y * rb.x), c1 = fma(rc.y, rb.y, -(rc.x * rb.x));
            sn = fma(s1, ra.y, c1 * ra.x);
            cs = fma(c1, ra.y, -(s1 * ra.x));
        } else {
            sincos(a, &sn, &cs);
        }
        const float af = (float)(a - kHalfPi);  // -de_ho (theta) / de_ve (phi), ray_trace.py:1529-1532
        a2 = af * af;
    };
    // theta columns of tile tx into buffer b (threads 0..15)
    auto fill_theta = [&](int tx, int b) {
        double sn, cs;
        float a2;
        fill_entry(tx * kTileW + tid, sn, cs, a2);
        s_sin_t[b][tid] = sn;
        s_cos_t[b][tid] = cs;
        // separable Gaussian feed taper a = A_theta[i] A_phi[j] (ray_trace.py:1573-1577)
        for (int f = 0; f < p.n_freq; ++f) s_amp_t[b][f][tid] = expf(g_freq.neg_c_fx[f] * a2);
    };

    const unsigned strips_per_rx = (unsigned)p.strips_x * (unsigned)p.tiles_y;
    double *const my_hist = s_hist + tid;
    constexpr int kHistStride = kBinThreads;
    // Strips are handed out dynamically (first one = blockIdx.x, then an atomic counter): their
    // cost varies with the position in the cone (binned / clipped / outside lens 1b's conic
    // domain, where warps fall back to fp32 seeds), and with only 10-100 strips per CTA a static
    // deal leaves SMs idle at the end (measured at 8 GPUs: 5.1 ms against 4.7 ms ideal).
    for (long long strip = blockIdx.x; strip < p.n_strips;) {
#ifdef SOSAT_STRIP_TIMING
        long long t_begin = 0;
        if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin)); }
#endif
        const int irx = (int)(strip / strips_per_rx);
        const unsigned srem = (unsigned)(strip - (long long)irx * strips_per_rx);
        // strips are taken from the rim of the cone inwards (rows and, within a row, columns
        // alternate between the two ends): the expensive, irregular strips -- clipped rays, fp32
        // fallback outside lens 1b's conic domain, bracketed solves -- go first and the uniform
        // central ones last, so the dynamic deal ends on short strips
        const int ky = (int)(srem / (unsigned)p.strips_x), kx = (int)(srem - (unsigned)ky * p.strips_x);
        const int ty_l = (ky & 1) ? p.tiles_y - 1 - (ky >> 1) : (ky >> 1);
        const int sx = (kx & 1) ? p.strips_x - 1 - (kx >> 1) : (kx >> 1);
        // a launch covers a block of consecutive launch rows or, with a row map, any set of tile
        // rows (multi-GPU: tile rows dealt cyclically, so every rank gets the same mix of rays)
        const int ty = p.tile_rows ? p.tile_rows[ty_l] : ty_l;
        const int tx_begin = sx * p.strip_len;
        const int tx_end = min(tx_begin + p.strip_len, p.tiles_x);
        if (irx != acc_rx) {
            if (acc_rx >= 0) flush_stats(acc_rx);
            acc_rx = irx;
        }
        __syncthreads();  // first pass: s_rot is complete (later: s_next has been read by everyone)
        if (tid < kTileW) {
            fill_theta(tx_begin, 0);
        } else if (tid < kTileW + kTileH) {
            const int l = tid - kTileW;
            double sn, cs;
            float a2;
            fill_entry((int)p.row_begin + ty * kTileH + l, sn, cs, a2);
            s_sin_p[l] = sn;
            s_cos_p[l] = cs;
            for (int f = 0; f < p.n_freq; ++f) s_amp_p[f][l] = expf(g_freq.neg_c_fy[f] * a2);
        } else if (tid < kTileW + kTileH + 3) {
            s_rx[tid - kTileW - kTileH] = d_rx[3 * irx + tid - kTileW - kTileH];
        }
        const int iph = ty * kTileH + lph;  // relative to row_begin
        unsigned okbits = 0u;               // bit 0 / 1: the lane's history of the last / last but one tile is finite
        unsigned strip_counts = 0u;         // valid (low 16 bits) and binned (high 16) rays of this warp in this strip
        int cur = 0;                        // s_hist[cur]: t(x-1), s_hist[cur ^ 1]: t(x-2) -> receives t(x)
        for (int tx = tx_begin, buf = 0; tx < tx_end; ++tx, buf ^= 1) {
            __syncthreads();  // tables of this tile complete; everybody is done with buffer buf ^ 1
            if (tid < kTileW && tx + 1 < tx_end) fill_theta(tx + 1, buf ^ 1);
            const double *rx = s_rx;  // feed position [mm], staged with the tables
            const double sth = s_sin_t[buf][lth], cth = s_cos_t[buf][lth];
            const double sph = s_sin_p[lph], cph = s_cos_p[lph];
            RayResult r;
            bool predicted = false;
            if (kCanPredict && p.predict) {
                const bool whole = tx * kTileW + kTileW <= p.n_scan;
                predicted = whole && __all_sync(0xffffffffu, okbits == 3u);
                double *h1 = my_hist + cur * (SOSAT_NUM_SURFACES * kHistStride);
                double *h2 = my_hist + (cur ^ 1) * (SOSAT_NUM_SURFACES * kHistStride);
                if (predicted)
                    trace_ray_predicted<GeoBuild<GEO>::generic, GeoBuild<GEO>::mask, GeoBuild<GEO>::edge>(
                        g_geom, rx[0], rx[1], rx[2], sth * cph, sth * sph, cth, r, h1, h2, kHistStride);
                else
                    trace_ray_newton_hist<N32, N64, GeoBuild<GEO>::generic, GeoBuild<GEO>::mask,
                                          GeoBuild<GEO>::edge, false>(
                        g_geom, rx[0], rx[1], rx[2], sth * cph, sth * sph, cth, r, h2, kHistStride);
                cur ^= 1;
                // a lane re-traced below leaves no usable history: its warp falls back for two tiles
                const bool fin = (r.opl - r.opl == 0.0) && !(r.flags & RAY_UNCONVERGED);
                okbits = ((okbits << 1) | (fin ? 1u : 0u)) & 3u;
            } else {
                trace_ray_newton<N32, N64, GeoBuild<GEO>::generic, GeoBuild<GEO>::mask,
                                 GeoBuild<GEO>::edge, false>(g_geom, rx[0], rx[1], rx[2], sth * cph,
                                                             sth * sph, cth, r);
            }
            const bool live = tx * kTileW + lth < p.n_scan && iph < p.n_rows;
            if (live && (r.flags & (RAY_UNCONVERGED | RAY_FLAGGED))) {
                // rare: the launch direction is re-read here (volatile) instead of being kept live
                // across the fast path, and results go through a local copy so that r itself
                // stays in registers.  (Queueing these rays for a second, densely packed pass was
                // built and measured: the first pass did not get faster -- the rare rays cluster
                // in few warps -- and the second pass cost 2.3 ms per 1e9 rays.)
                const double sth2 = *(volatile double *)&s_sin_t[buf][lth], cth2 = *(volatile double *)&s_cos_t[buf][lth];
                const double sph2 = *(volatile double *)&s_sin_p[lph], cph2 = *(volatile double *)&s_cos_p[lph];
                RayResult rs;
                rs.valid = r.valid;
                rs.flags = r.flags;
                bool redone = false;
                if (r.flags & RAY_UNCONVERGED) {  // adaptive Newton refinement, out of line
                    trace_ray_adaptive<GeoBuild<GEO>::generic, GeoBuild<GEO>::mask, GeoBuild<GEO>::edge>(
                        g_geom, rx[0], rx[1], rx[2], sth2 * cph2, sth2 * sph2, cth2, rs);
                    redone = true;
                }
                if (rs.valid && (rs.flags & RAY_FLAGGED)) {
                    // < 0.15 % of rays: the reference's bracketed solve
                    trace_ray_bracketed(g_geom, rx[0], rx[1], rx[2], sth2 * cph2, sth2 * sph2, cth2, rs);
                    redone = true;
                    atomicAdd(&s_rare[0], 1u);
                    if (rs.flags & RAY_BRACKET_FAIL) atomicAdd(&s_rare[1], 1u);
                }
                if (redone) {
                    r.ax = rs.ax; r.az = rs.az; r.opl = rs.opl;
                    r.dx = rs.dx; r.dy = rs.dy; r.dz = rs.dz;
                    r.valid = rs.valid;
                }
            }
            const bool valid = live && r.valid;
            int bin = -1;
            if (valid) {
                const int ix = __double2int_rd((r.ax + p.half_extent) * p.inv_cell);
                const int iz = __double2int_rd((r.az + p.half_extent) * p.inv_cell);
                if ((unsigned)ix < (unsigned)p.grid_n && (unsigned)iz < (unsigned)p.grid_n)
                    bin = iz * p.grid_n + ix;
                s_sum[0][tid] += r.opl;
                s_sum[1][tid] += r.dx;
                s_sum[2][tid] += r.dy;
                s_sum[3][tid] += r.dz;
            }
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const unsigned bmask = __ballot_sync(0xffffffffu, bin >= 0);
            strip_counts += (unsigned)__popc(vmask) | ((unsigned)__popc(bmask) << 16);  // <= 64 x 32 per strip
            if (re == nullptr || bmask == 0u) continue;

            // ---- run-length aggregation along the lane order --------------------------------
            const int prev = __shfl_up_sync(0xffffffffu, bin, 1);
            const bool head = lane == 0 || prev != bin;
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            // bit b of m: lane + 1 + b starts a new run; a virtual head sits just past lane 31
            const unsigned m = ((heads >> lane) >> 1) | (1u << (31 - lane));
            const int rest = __ffs(m) - 1;  // lanes lane+1 .. lane+rest continue this lane's run
            const bool flush = head && bin >= 0;
            if (flush) atomicAdd(cnt + (long long)irx * plane + bin, (float)(rest + 1));

            const double dopl = r.opl - p.opl_ref;
            float *re_f = re + (long long)irx * p.n_freq * plane;
            float *im_f = im + (long long)irx * p.n_freq * plane;
            auto emit = [&](int f, float *re_p, float *im_p) {
                float vre = 0.f, vim = 0.f;
                if (bin >= 0) {
                    const float a = s_amp_t[buf][f][lth] * s_amp_p[f][lph];
                    double cyc = dopl * g_freq.kk[f];
                    // phase reduced to [-1/2, 1/2] cycles in fp64: cyc - rint(cyc) with the
                    // 1.5 * 2^52 trick (two DADDs; an FRND.F64 would go through the XU pipe)
                    cyc -= __dadd_rn(__dadd_rn(cyc, 6755399441055744.0), -6755399441055744.0);
                    float sn, cs;
                    __sincosf(6.2831853071795865f * (float)cyc, &sn, &cs);  // SFU, |err| < 1e-6
                    vre = a * cs;
                    vim = a * sn;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float ore = __shfl_down_sync(0xffffffffu, vre, o);
                    const float oim = __shfl_down_sync(0xffffffffu, vim, o);
                    if (o <= rest) {
                        vre += ore;
                        vim += oim;
                    }
                }
                if (flush) {
                    atomicAdd(re_p + bin, vre);
                    atomicAdd(im_p + bin, vim);
                }
            };
            if (p.n_freq == 1) {  // the common case without the loop's bookkeeping
                emit(0, re_f, im_f);
            } else {
                for (int f = 0; f < p.n_freq; ++f, re_f += plane, im_f += plane) emit(f, re_f, im_f);
            }
        }
#ifdef SOSAT_STRIP_TIMING
        if (tid == 0 && strip < (1 << 17)) {
            long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_strip_t0[strip] = t_begin; g_strip_t1[strip] = t_end; g_strip_sm[strip] = (int)smid;
        }
#endif
        if (lane == 0) {
            s_nvalid[warp] += strip_counts & 0xffffu;
            s_nbinned[warp] += strip_counts >> 16;
        }
        if (tid == 0) s_next = (long long)gridDim.x + (long long)atomicAdd(&g_next_strip, 1ull);
        __syncthreads();  // also: every warp is done with this strip's tables
        strip = s_next;
    }
    if (acc_rx >= 0) flush_stats(acc_rx);
}

// ---- getNearField post-processing (ray_trace.py:1599-1614) ---------------------------------
__global__ void nf_sums_kernel(const double *__restrict__ out, long long n,
                               double *__restrict__ work) {
    double acc[5] = {0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        if (out[4 * n + i] != 0.0) {  // xx = np.where(out[4] != 0)
            acc[0] += 1.0;
            acc[1] += out[3 * n + i];
            acc[2] += out[8 * n + i];
            acc[3] += out[9 * n + i];
            acc[4] += out[10 * n + i];
        }
    }
    block_accumulate<5>(acc, work);
}

// numpy.mod for floats: result takes the sign of the divisor
__device__ __forceinline__ double np_mod(double x, double y) {
    double r = fmod(x, y);
    if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r += y;
    return r;
}

__global__ void nf_phase_kernel(const double *__restrict__ out, long long n, double k,
                                double *__restrict__ xyap, uint8_t *__restrict__ mask,
                                double *__restrict__ work) {
    const double mean_opl = work[1] / work[0];
    double acc[1] = {0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double a = out[4 * n + i];
        const bool m = a != 0.0;
        double p = 0.0;
        if (m) {
            // np.mod(k * (OPL - mean) / 1e3 / 2, 2 pi), ray_trace.py:1601
            p = np_mod(__ddiv_rn(__ddiv_rn(__dmul_rn(k, out[3 * n + i] - mean_opl), 1e3), 2.0),
                       6.283185307179586);
            acc[0] += p;
        }
        xyap[0 * n + i] = out[0 * n + i] / 1e1;  // x_sim [cm]
        xyap[1 * n + i] = out[2 * n + i] / 1e1;  // y_sim [cm]
        xyap[2 * n + i] = a;
        xyap[3 * n + i] = p;
        mask[i] = m ? 1 : 0;
    }
    block_accumulate<1>(acc, work + 5);
}

__global__ void nf_finish_kernel(long long n, double *__restrict__ xyap,
                                 const uint8_t *__restrict__ mask, double *__restrict__ tan_og,
                                 const double *__restrict__ work) {
    const double mean_p = work[5] / work[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        if (mask[i]) xyap[3 * n + i] -= mean_p;  // p_sim - np.mean(p_sim), ray_trace.py:1614
    if (blockIdx.x == 0 && threadIdx.x < 3) tan_og[threadIdx.x] = work[2 + threadIdx.x] / work[0];
}

__global__ void bins_to_field_kernel(const float *__restrict__ re, const float *__restrict__ im,
                                     const float *__restrict__ cnt, long long n, float cs,
                                     float sn, float2 *__restrict__ b) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float c = cnt[i];
        float2 v = make_float2(0.f, 0.f);
        if (c > 0.f) {
            const float inv = 1.0f / c, x = re[i] * inv, y = im[i] * inv;
            v.x = x * cs + y * sn;  // (x + i y) * exp(-i phi0)
            v.y = y * cs - x * sn;
        }
        b[i] = v;
    }
}

// Same normalisation with the phase reference taken on the device from the statistics row of the
// launch (mean OPL of the valid rays): no host round trip between the trace and the far field.
__global__ void bins_to_field_stats_kernel(const float *__restrict__ re, const float *__restrict__ im,
                                           const float *__restrict__ cnt, long long n,
                                           const double *__restrict__ stats, double k_over_2e3,
                                           double opl_ref, float2 *__restrict__ b) {
    const double nv = stats[SOSAT_STAT_N_VALID];
    const double mean_opl = nv > 0 ? stats[SOSAT_STAT_SUM_OPL] / nv : 0.0;
    double sn, cs;
    sincos(k_over_2e3 * (mean_opl - opl_ref), &sn, &cs);
    const float fcs = (float)cs, fsn = (float)sn;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float c = cnt[i];
        float2 v = make_float2(0.f, 0.f);
        if (c > 0.f) {
            const float inv = 1.0f / c, x = re[i] * inv, y = im[i] * inv;
            v.x = x * fcs + y * fsn;  // (x + i y) * exp(-i phi0)
            v.y = y * fcs - x * fsn;
        }
        b[i] = v;
    }
}

// The whole stack of (receiver, frequency) fields of a sweep in ONE launch: blockIdx.y = field
// irx * n_freq + f; the statistics row is the receiver's, the phase scale the frequency's.
struct FieldScales {
    double k_over_2e3[SOSAT_MAX_FREQS];
};
__global__ void bins_to_field_stats_batch_kernel(const float *__restrict__ re, const float *__restrict__ im,
                                                 const float *__restrict__ cnt, long long n, int n_freq,
                                                 const double *__restrict__ stats, FieldScales ks,
                                                 double opl_ref, float2 *__restrict__ b) {
    const int fld = blockIdx.y, irx = fld / n_freq, f = fld - irx * n_freq;
    const double *st = stats + (long long)irx * SOSAT_NUM_STATS;
    const double nv = st[SOSAT_STAT_N_VALID];
    const double mean_opl = nv > 0 ? st[SOSAT_STAT_SUM_OPL] / nv : 0.0;
    double sn, cs;
    sincos(ks.k_over_2e3[f] * (mean_opl - opl_ref), &sn, &cs);
    const float fcs = (float)cs, fsn = (float)sn;
    re += (long long)fld * n;
    im += (long long)fld * n;
    cnt += (long long)irx * n;
    b += (long long)fld * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float c = cnt[i];
        float2 v = make_float2(0.f, 0.f);
        if (c > 0.f) {
            const float inv = 1.0f / c, x = re[i] * inv, y = im[i] * inv;
            v.x = x * fcs + y * fsn;  // (x + i y) * exp(-i phi0)
            v.y = y * fcs - x * fsn;
        }
        b[i] = v;
    }
}

// ---- FP64 FMA throughput microkernel (roofline denominator of K1) ---------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

// ---- host side ---------------------------------------------------------------------------------
// Newton schedule (fp32 steps, fp64 steps) of the fast path.  Default "2 + 1": one Newton and
// one Halley step in fp32, ONE fp64 Newton step, slope factor carried to the new root to first
// order (trace_device.cuh), plus the adaptive refinement for lanes whose fp64 step is too long.
// SOSAT_TRACE_VARIANT = 0: 0+5, 2: 3+2, 6: 2+2 (the default until v3) select the other
// compiled schedules for experiments and for the parity test of the schedules.
static int trace_variant() {
#ifdef SOSAT_EXPERIMENTS
    const char *e = getenv("SOSAT_TRACE_VARIANT");
    return e ? atoi(e) : 7;
#else
    return 7;  // the other schedules are compiled with -DSOSAT_EXPERIMENTS only
#endif
}

static void fill_linspace(double cone, int n_scan, double &start, double &stop, double &step) {
    start = kHalfPi - cone;
    stop = kHalfPi + cone;
    step = n_scan > 1 ? (stop - start) / (double)(n_scan - 1) : 0.0;
}

static sosat_geom_t g_host_geom[64];

static int current_device(int *dev) {
    SOSAT_CUDA_OK(cudaGetDevice(dev));
    if (*dev < 0 || *dev >= 64) return set_error(SOSAT_ERR_UNSUPPORTED, "device index %d", *dev);
    return 0;
}

}  // namespace sosat

using namespace sosat;

extern "C" int sosat_set_geometry(const sosat_geom_t *h, void *stream) {
    SOSAT_REQUIRE(h != nullptr, "sosat_set_geometry: null geometry");
    int dev;
    if (int rc = current_device(&dev)) return rc;
    GeomC g;
    memset(&g, 0, sizeof g);
    bool generic = false;
    int refine_mask = 0, edge_mask = 0;
    for (int i = 0; i < SOSAT_NUM_SURFACES; ++i) {
        const sosat_surface_t &s = h->surf[i];
        SOSAT_REQUIRE(s.n_in > 0 && s.n_out > 0, "surface %d: refractive index <= 0", i);
        SurfC &c = g.s[i];
        c.Kp = 0.01 * (1.0 + s.k) * s.c * s.c;
        c.cp = 0.1 * s.c;
        c.a1p = 0.1 * s.a1;
        c.a2p = 1e-3 * s.a2;
        c.a3p = 1e-5 * s.a3;
        c.a1p2 = 2.0 * c.a1p;
        c.a2p4 = 4.0 * c.a2p;
        c.a3p6 = 6.0 * c.a3p;
        c.y0 = s.y0;
        c.mu = s.n_in / s.n_out;
        c.mu2 = c.mu * c.mu;
        c.n_in = s.n_in;
        c.t_lo = s.t_lo;
        c.t_hi = s.t_hi;
        c.omm2 = 1.0 - c.mu2;
        c.cKh = 0.5 * c.cp * c.Kp;
        c.cK = c.Kp != 0.0 ? c.cp / c.Kp : 0.0;
        c.y0K = c.y0 - c.cK;
        c.refine = fabs(c.cK) > 300.0;
        refine_mask |= c.refine << i;
        edge_mask |= (c.Kp > 0.0) << i;
        if (c.Kp == 0.0 ? c.cp != 0.0 : !(fabs(c.cK) <= 1e6)) generic = true;
        SurfF &f = g.f[i];
        f.Kp = (float)c.Kp; f.cp = (float)c.cp; f.a1p = (float)c.a1p; f.a2p = (float)c.a2p;
        f.a3p = (float)c.a3p; f.a1p2 = (float)c.a1p2; f.a2p4 = (float)c.a2p4;
        f.a3p6 = (float)c.a3p6;
        f.cK = (float)c.cK;
        f.cKh = (float)c.cKh;
    }
    g.y_lyot = h->y_lyot;
    g.y_source = h->y_source;
    g.clip_l3_sq = h->clip_l3 * h->clip_l3;
    g.clip_stop_sq = h->clip_stop * h->clip_stop;
    g.cone = h->cone;
    SOSAT_CUDA_OK(cudaMemcpyToSymbolAsync(g_geom, &g, sizeof g, 0, cudaMemcpyHostToDevice,
                                          (cudaStream_t)stream));
    // the staging copy above reads &g before returning only for pageable memory semantics
    SOSAT_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    g_host_geom[dev] = *h;
    g_geom_set[dev] = true;
    const char *force = getenv("SOSAT_TRACE_GEO");  // testing hook: force build 1 or 2
    int build = generic ? 2 : (refine_mask == 0x04 && edge_mask == 0x10 ? 0 : 1);
    if (force && atoi(force) > build) build = atoi(force) > 2 ? 2 : atoi(force);
    g_geom_build[dev] = build;
    return SOSAT_OK;
}

template <int N32, int N64, bool COMPACT>
static void launch_rows(int geo, const RowsParams &p, double *out, float4 *samples, double *stats,
                        cudaStream_t st) {
    const int threads = 256;
    const unsigned blocks = (unsigned)((p.n + threads - 1) / threads);
    switch (geo) {
        case 0: trace_rows_kernel<N32, N64, 0, COMPACT><<<blocks, threads, 0, st>>>(p, out, samples, stats); break;
        case 1: trace_rows_kernel<N32, N64, 1, COMPACT><<<blocks, threads, 0, st>>>(p, out, samples, stats); break;
        default: trace_rows_kernel<N32, N64, 2, COMPACT><<<blocks, threads, 0, st>>>(p, out, samples, stats); break;
    }
}

static int trace_rows_common(const double *h_rx, const sosat_freq_t *h_freq, int n_scan,
                             int64_t ray_begin, int64_t ray_end, double opl_ref, double *d_out,
                             float4 *d_samples, double *d_stats, void *stream) {
    int dev;
    if (int rc = current_device(&dev)) return rc;
    if (!g_geom_set[dev]) return set_error(SOSAT_ERR_NO_GEOMETRY, "call sosat_set_geometry first");
    SOSAT_REQUIRE(h_rx && h_freq && (d_out || d_samples), "null pointer");
    SOSAT_REQUIRE(n_scan >= 1, "n_scan must be >= 1 (got %d)", n_scan);
    const int64_t total = (int64_t)n_scan * n_scan;
    SOSAT_REQUIRE(ray_begin >= 0 && ray_end >= ray_begin && ray_end <= total,
                  "ray range [%lld, %lld) outside [0, %lld]", (long long)ray_begin,
                  (long long)ray_end, (long long)total);
    SOSAT_REQUIRE(h_freq->fwhp_x != 0 && h_freq->fwhp_y != 0, "feed FWHP must be non-zero");
    RowsParams p;
    p.rx[0] = h_rx[0]; p.rx[1] = h_rx[1]; p.rx[2] = h_rx[2];
    p.fwhp_x = h_freq->fwhp_x;
    p.fwhp_y = h_freq->fwhp_y;
    p.opl_ref = opl_ref;
    p.kk = h_freq->k_over_2e3 / 6.283185307179586;
    fill_linspace(g_host_geom[dev].cone, n_scan, p.lin_start, p.lin_stop, p.lin_step);
    p.n_scan = n_scan;
    p.ray_begin = ray_begin;
    p.n = ray_end - ray_begin;
    if (p.n == 0) return SOSAT_OK;
    SOSAT_REQUIRE((p.n + 255) / 256 < 2147483647LL, "too many rays for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const int geo = g_geom_build[dev];
    if (d_samples) {
        launch_rows<2, 1, true>(geo, p, nullptr, d_samples, d_stats, st);
    } else {
        switch (trace_variant()) {
#ifdef SOSAT_EXPERIMENTS
            case 0: launch_rows<0, 5, false>(geo, p, d_out, nullptr, d_stats, st); break;
            case 2: launch_rows<3, 2, false>(geo, p, d_out, nullptr, d_stats, st); break;
            case 6: launch_rows<2, 2, false>(geo, p, d_out, nullptr, d_stats, st); break;
#endif
            default: launch_rows<2, 1, false>(geo, p, d_out, nullptr, d_stats, st); break;
        }
    }
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_trace_rays(const double *h_rx, const sosat_freq_t *h_freq, int n_scan,
                                int64_t ray_begin, int64_t ray_end, double *d_out,
                                double *d_stats, void *stream) {
    SOSAT_REQUIRE(d_out, "sosat_trace_rays: null pointer");
    return trace_rows_common(h_rx, h_freq, n_scan, ray_begin, ray_end, 0.0, d_out, nullptr, d_stats,
                             stream);
}

extern "C" int sosat_trace_samples(const double *h_rx, const sosat_freq_t *h_freq, int n_scan,
                                   int64_t ray_begin, int64_t ray_end, double opl_ref,
                                   float *d_xyb, double *d_stats, void *stream) {
    SOSAT_REQUIRE(d_xyb, "sosat_trace_samples: null pointer");
    SOSAT_REQUIRE(((uintptr_t)d_xyb & 15) == 0, "sosat_trace_samples: d_xyb must be 16-byte aligned");
    return trace_rows_common(h_rx, h_freq, n_scan, ray_begin, ray_end, opl_ref, nullptr,
                             (float4 *)d_xyb, d_stats, stream);
}

extern "C" int sosat_rx_to_lyot_host(const sosat_geom_t *h_geom, const double *h_rx,
                                     const sosat_freq_t *h_freq, int n_scan, int64_t ray_begin,
                                     int64_t ray_end, double *h_out, double *h_stats) {
    SOSAT_REQUIRE(h_geom && h_rx && h_freq && h_out, "sosat_rx_to_lyot_host: null pointer");
    SOSAT_REQUIRE(n_scan >= 1, "n_scan must be >= 1 (got %d)", n_scan);
    if (int rc = sosat_set_geometry(h_geom, nullptr)) return rc;
    const int64_t n = ray_end - ray_begin;
    SOSAT_REQUIRE(n >= 0 && ray_begin >= 0 && ray_end <= (int64_t)n_scan * n_scan,
                  "ray range [%lld, %lld) outside [0, %lld]", (long long)ray_begin,
                  (long long)ray_end, (long long)n_scan * n_scan);
    if (n == 0) return SOSAT_OK;
    double *d_out = nullptr, *d_stats = nullptr;
    SOSAT_CUDA_OK(cudaMalloc(&d_out, sizeof(double) * SOSAT_OUT_ROWS * n));
    cudaError_t e = cudaMalloc(&d_stats, sizeof(double) * SOSAT_NUM_STATS);
    if (e != cudaSuccess) {
        cudaFree(d_out);
        return set_error(SOSAT_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    cudaMemset(d_stats, 0, sizeof(double) * SOSAT_NUM_STATS);
    int rc = sosat_trace_rays(h_rx, h_freq, n_scan, ray_begin, ray_end, d_out, d_stats, nullptr);
    if (rc == 0) {
        e = cudaMemcpy(h_out, d_out, sizeof(double) * SOSAT_OUT_ROWS * n, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && h_stats)
            e = cudaMemcpy(h_stats, d_stats, sizeof(double) * SOSAT_NUM_STATS,
                           cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_error(SOSAT_ERR_CUDA, "cudaMemcpy: %s", cudaGetErrorString(e));
    }
    cudaFree(d_out);
    cudaFree(d_stats);
    return rc;
}

// Launch the persistent binning kernel with grid = SMs x resident CTAs per SM.  The strip length
// is chosen here because it needs the grid size: strips are handed out dynamically, so there must
// be many more strips than CTAs; a predicted-seed strip pays two fp32-seeded tiles at its start.
template <int N32, int N64, int MINB, int GEO>
static void launch_bin(BinParams p, const double *rx, float *re, float *im, float *cnt,
                       double *stats, cudaStream_t st) {
    auto kern = trace_bin_kernel<N32, N64, MINB, GEO>;
    const size_t smem = p.predict ? (size_t)kHistBytes : 0;
    // per device and per (kernel, smem) variant, queried once: attributes, occupancy, the address
    // of the work-queue counter (host-side latency in front of small launches)
    static bool ready[64][2] = {};
    static int occ_cache[64][2] = {};
    static void *counter_cache[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    const int v = p.predict ? 1 : 0;
    if (!ready[dev][v]) {
        // MINB CTAs of ~45 KB each: ask for the shared-memory-heavy split of the L1 / shared array
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        // static tables (25 KB) + history (24 KB) exceed the 48 KB default limit
        if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache[dev][v], kern, kBinThreads, smem);
        cudaGetSymbolAddress(&counter_cache[dev], g_next_strip);
        ready[dev][v] = true;
    }
    const int occ = occ_cache[dev][v];
    const long long want = (long long)num_sms() * (occ > 0 ? occ : 1);
    const long long n_tiles = (long long)p.tiles_x * p.tiles_y * p.n_rx;
    long long sl = 64;
    if (const char *e = getenv("SOSAT_STRIP_LEN")) {  // testing / tuning hook
        sl = atoll(e);
    } else {
        // >= 8 strips per CTA: with the rim-first order the dynamic deal ends on the uniform
        // central strips, so few, long strips are enough (1/8 of the benchmark bundle: 4.29 ms
        // with 13 strips of 64 tiles per CTA, 4.36 ms with 33 strips of 25), and every strip
        // start costs two fp32-seeded tiles
        if (sl > n_tiles / (8 * want)) sl = n_tiles / (8 * want);
        if (sl < 8) p.predict = 0;  // short strips: the two seeding tiles would dominate
    }
    if (sl < 1) sl = 1;
    p.strip_len = (int)sl;
    p.strips_x = (p.tiles_x + p.strip_len - 1) / p.strip_len;
    p.n_strips = (long long)p.strips_x * p.tiles_y * p.n_rx;
    const int blocks = (int)(p.n_strips < want ? p.n_strips : want);
    cudaMemsetAsync(counter_cache[dev], 0, sizeof(unsigned long long), st);
    kern<<<blocks, kBinThreads, smem, st>>>(p, rx, re, im, cnt, stats);
}

template <int N32, int N64>
static void launch_bin_minb(int geo, int minb, const BinParams &p, const double *rx,
                            float *re, float *im, float *cnt, double *stats, cudaStream_t st) {
    if (geo == 2) return launch_bin<N32, N64, 4, 2>(p, rx, re, im, cnt, stats, st);
    if (geo == 1) return launch_bin<N32, N64, 4, 1>(p, rx, re, im, cnt, stats, st);
    switch (minb) {
#ifdef SOSAT_EXPERIMENTS
        case 3: launch_bin<N32, N64, 3, 0>(p, rx, re, im, cnt, stats, st); break;
#endif
        default: launch_bin<N32, N64, 4, 0>(p, rx, re, im, cnt, stats, st); break;
    }
}

// resident CTAs per SM the binning kernel is compiled for (register budget 80 / 64)
static int trace_minb() {
#ifdef SOSAT_EXPERIMENTS
    const char *e = getenv("SOSAT_TRACE_MINB");
    return e ? atoi(e) : 4;
#else
    return 4;
#endif
}

static int trace_bin_common(const double *d_rx, int n_rx, const sosat_freq_t *h_freq,
                            int n_freq, int n_scan, int64_t row_begin, int64_t row_end,
                            const int *d_tile_rows, int n_tile_rows, int grid_n, double extent_mm,
                            double opl_ref, float *d_re, float *d_im, float *d_cnt, double *d_stats,
                            void *stream) {
    int dev;
    if (int rc = current_device(&dev)) return rc;
    if (!g_geom_set[dev]) return set_error(SOSAT_ERR_NO_GEOMETRY, "call sosat_set_geometry first");
    SOSAT_REQUIRE(d_rx && h_freq, "sosat_trace_bin: null pointer");
    SOSAT_REQUIRE(n_rx >= 1, "n_rx must be >= 1");
    SOSAT_REQUIRE(n_freq >= 1 && n_freq <= SOSAT_MAX_FREQS, "n_freq must be in [1, %d]",
                  SOSAT_MAX_FREQS);
    SOSAT_REQUIRE(n_scan >= 1, "n_scan must be >= 1");
    SOSAT_REQUIRE(row_begin >= 0 && row_end >= row_begin && row_end <= n_scan,
                  "row range [%lld, %lld) outside [0, %d]", (long long)row_begin,
                  (long long)row_end, n_scan);
    const bool binning = d_re || d_im || d_cnt;
    if (binning) {
        SOSAT_REQUIRE(d_re && d_im && d_cnt, "d_re, d_im, d_cnt must be all set or all NULL");
        SOSAT_REQUIRE(grid_n >= 1 && grid_n <= 32768, "grid_n out of range");
        SOSAT_REQUIRE(extent_mm > 0, "extent_mm must be positive");
    }
    if (row_end == row_begin) return SOSAT_OK;
    BinFreq bf;
    memset(&bf, 0, sizeof bf);
    for (int f = 0; f < n_freq; ++f) {
        SOSAT_REQUIRE(h_freq[f].fwhp_x != 0 && h_freq[f].fwhp_y != 0, "feed FWHP must be non-zero");
        bf.neg_c_fx[f] = (float)(-kFourLn2 / (h_freq[f].fwhp_x * h_freq[f].fwhp_x));
        bf.neg_c_fy[f] = (float)(-kFourLn2 / (h_freq[f].fwhp_y * h_freq[f].fwhp_y));
        bf.kk[f] = h_freq[f].k_over_2e3 / 6.283185307179586;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SOSAT_CUDA_OK(cudaMemcpyToSymbolAsync(g_freq, &bf, sizeof bf, 0, cudaMemcpyHostToDevice, st));
    BinParams p;
    const int minb = trace_minb();
    const int geo = g_geom_build[dev];
    fill_linspace(g_host_geom[dev].cone, n_scan, p.lin_start, p.lin_stop, p.lin_step);
    p.opl_ref = opl_ref;
    p.half_extent = binning ? 0.5 * extent_mm : 0.0;
    p.inv_cell = binning ? (double)grid_n / extent_mm : 0.0;
    p.row_begin = row_begin;
    p.n_scan = n_scan;
    p.n_rows = (int)(row_end - row_begin);
    p.n_rx = n_rx;
    p.n_freq = n_freq;
    p.grid_n = binning ? grid_n : 0;
    p.tiles_x = (n_scan + kTileW - 1) / kTileW;
    p.tiles_y = d_tile_rows ? n_tile_rows : (p.n_rows + kTileH - 1) / kTileH;
    p.tile_rows = d_tile_rows;
    SOSAT_REQUIRE((long long)p.tiles_x * p.tiles_y * n_rx < 2147483647LL,
                  "too many ray tiles for one launch (%lld): split the receiver list or the row range",
                  (long long)p.tiles_x * p.tiles_y * n_rx);
    // Predicted seeds (trace_device.cuh: trace_ray_predicted): the extrapolation error over one
    // tile step is (16 dtheta)^2 |t''| <= 2.4e3 mm/rad^2 x (16 step)^2: 4e-4 mm at n_scan = 31 623,
    // 1.5e-3 mm at 16 384.  A lane whose fp64 step exceeds 2^-10 mm is re-traced adaptively and its
    // warp falls back to the fp32 seeds for two tiles, so a grid that is too coarse for the
    // prediction costs nothing but the attempt.  Measured (scripts/ab_predict_threshold.py, time
    // against the fp32 seeds, bin counts identical everywhere): 0.74 at n_scan >= 26 000, 0.78 at
    // 16 384, 0.85 at 12 000, 0.91 at 10 000, 0.99-1.01 at <= 8 192.  Enabled for
    // 16 step <= 1.3e-3 rad (n_scan >= ~9 850 for the 0.8 rad cone); SOSAT_TRACE_PREDICT=0/1 forces it.
    p.predict = 16.0 * p.lin_step <= 1.3e-3 && p.lin_step > 0.0;
    if (const char *e = getenv("SOSAT_TRACE_PREDICT")) p.predict = atoi(e) != 0;
    if (trace_variant() != 7 || geo == 2) p.predict = 0;
    switch (trace_variant()) {
#ifdef SOSAT_EXPERIMENTS
        case 0: launch_bin_minb<0, 5>(geo, minb, p, d_rx, d_re, d_im, d_cnt, d_stats, st); break;
        case 2: launch_bin_minb<3, 2>(geo, minb, p, d_rx, d_re, d_im, d_cnt, d_stats, st); break;
        case 6: launch_bin_minb<2, 2>(geo, minb, p, d_rx, d_re, d_im, d_cnt, d_stats, st); break;
#endif
        default: launch_bin_minb<2, 1>(geo, minb, p, d_rx, d_re, d_im, d_cnt, d_stats, st); break;
    }
    SOSAT_CUDA_OK(cudaGetLastError());
    // bf lives on this stack frame: the async symbol copy from pageable memory is staged
    // before the call returns, so no synchronisation is needed here.
    return SOSAT_OK;
}

extern "C" int sosat_trace_bin(const double *d_rx, int n_rx, const sosat_freq_t *h_freq,
                               int n_freq, int n_scan, int64_t row_begin, int64_t row_end,
                               int grid_n, double extent_mm, double opl_ref, float *d_re,
                               float *d_im, float *d_cnt, double *d_stats, void *stream) {
    return trace_bin_common(d_rx, n_rx, h_freq, n_freq, n_scan, row_begin, row_end, nullptr, 0, grid_n,
                            extent_mm, opl_ref, d_re, d_im, d_cnt, d_stats, stream);
}

extern "C" int sosat_trace_bin_tiles(const double *d_rx, int n_rx, const sosat_freq_t *h_freq,
                                     int n_freq, int n_scan, const int *d_tile_rows,
                                     int n_tile_rows, int grid_n, double extent_mm, double opl_ref,
                                     float *d_re, float *d_im, float *d_cnt, double *d_stats,
                                     void *stream) {
    SOSAT_REQUIRE(n_tile_rows >= 0, "n_tile_rows must be >= 0");
    SOSAT_REQUIRE(n_tile_rows == 0 || d_tile_rows, "sosat_trace_bin_tiles: null tile row list");
    if (n_tile_rows == 0) return SOSAT_OK;
    return trace_bin_common(d_rx, n_rx, h_freq, n_freq, n_scan, 0, n_scan, d_tile_rows, n_tile_rows,
                            grid_n, extent_mm, opl_ref, d_re, d_im, d_cnt, d_stats, stream);
}

#ifdef SOSAT_STRIP_TIMING
extern "C" int sosat_debug_strip_times(long long *h_t0, long long *h_t1, int *h_sm, int n) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h_t0, g_strip_t0, sizeof(long long) * n);
    cudaMemcpyFromSymbol(h_t1, g_strip_t1, sizeof(long long) * n);
    cudaMemcpyFromSymbol(h_sm, g_strip_sm, sizeof(int) * n);
    return 0;
}
#endif

extern "C" int sosat_near_field_arrays(const double *d_out, int64_t n, double k, double *d_xyap,
                                       uint8_t *d_mask, double *d_tan_og, double *d_work,
                                       void *stream) {
    SOSAT_REQUIRE(d_out && d_xyap && d_mask && d_tan_og && d_work,
                  "sosat_near_field_arrays: null pointer");
    SOSAT_REQUIRE(n >= 0, "negative ray count");
    cudaStream_t st = (cudaStream_t)stream;
    SOSAT_CUDA_OK(cudaMemsetAsync(d_work, 0, sizeof(double) * SOSAT_NUM_STATS, st));
    if (n == 0) return SOSAT_OK;
    const int threads = 256;
    long long blocks = (n + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    nf_sums_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_out, n, d_work);
    nf_phase_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_out, n, k, d_xyap, d_mask, d_work);
    nf_finish_kernel<<<(unsigned)blocks, threads, 0, st>>>(n, d_xyap, d_mask, d_tan_og, d_work);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_bins_to_field(const float *d_re, const float *d_im, const float *d_cnt,
                                   int64_t n, double phase_offset_rad, float *d_b, void *stream) {
    SOSAT_REQUIRE(d_re && d_im && d_cnt && d_b, "sosat_bins_to_field: null pointer");
    SOSAT_REQUIRE(n >= 0, "negative cell count");
    if (n == 0) return SOSAT_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    bins_to_field_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_re, d_im, d_cnt, n, (float)cos(phase_offset_rad), (float)sin(phase_offset_rad),
        (float2 *)d_b);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_bins_to_field_stats(const float *d_re, const float *d_im, const float *d_cnt,
                                         int64_t n, const double *d_stats, double k_over_2e3,
                                         double opl_ref, float *d_b, void *stream) {
    SOSAT_REQUIRE(d_re && d_im && d_cnt && d_stats && d_b, "sosat_bins_to_field_stats: null pointer");
    SOSAT_REQUIRE(n >= 0, "negative cell count");
    if (n == 0) return SOSAT_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    bins_to_field_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_re, d_im, d_cnt, n, d_stats, k_over_2e3, opl_ref, (float2 *)d_b);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_bins_to_field_stats_batch(const float *d_re, const float *d_im,
                                               const float *d_cnt, int64_t n, int n_rx, int n_freq,
                                               const double *d_stats, const double *h_k_over_2e3,
                                               double opl_ref, float *d_b, void *stream) {
    SOSAT_REQUIRE(d_re && d_im && d_cnt && d_stats && d_b && h_k_over_2e3,
                  "sosat_bins_to_field_stats_batch: null pointer");
    SOSAT_REQUIRE(n >= 0, "negative cell count");
    SOSAT_REQUIRE(n_rx >= 0 && n_freq >= 1 && n_freq <= SOSAT_MAX_FREQS &&
                      (long long)n_rx * n_freq <= 65535,
                  "n_rx x n_freq = %d x %d out of range", n_rx, n_freq);
    if (n == 0 || n_rx == 0) return SOSAT_OK;
    FieldScales ks;
    memset(&ks, 0, sizeof ks);
    for (int f = 0; f < n_freq; ++f) ks.k_over_2e3[f] = h_k_over_2e3[f];
    const int fields = n_rx * n_freq;
    long long blocks = (n + 255) / 256;
    long long cap = (long long)num_sms() * 8 / fields;
    if (cap < 4) cap = 4;
    if (blocks > cap) blocks = cap;
    bins_to_field_stats_batch_kernel<<<dim3((unsigned)blocks, (unsigned)fields), 256, 0,
                                       (cudaStream_t)stream>>>(d_re, d_im, d_cnt, n, n_freq, d_stats,
                                                               ks, opl_ref, (float2 *)d_b);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_measure_fp64_peak(double *flops_per_s, double *ms_out) {
    SOSAT_REQUIRE(flops_per_s, "sosat_measure_fp64_peak: null pointer");
    double *d = nullptr;
    SOSAT_CUDA_OK(cudaMalloc(&d, 64));
    const int iters = 1 << 15, threads = 256, blocks = num_sms() * 16;
    cudaEvent_t e0, e1;
    SOSAT_CUDA_OK(cudaEventCreate(&e0));
    SOSAT_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        SOSAT_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    SOSAT_CUDA_OK(cudaGetLastError());
    const double fl = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
    *flops_per_s = fl / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return SOSAT_OK;
}
