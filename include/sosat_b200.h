/*
 * sosat_b200.h -- C-ABI of libsosat_b200.so: the B200 (sm_100a) implementation of the
 * sosat_optics beam-simulation hot path (ray trace feed -> source plane, near-field
 * binning, far-field transform).
 *
 * The reference (McMahonCosmologyGroup/sosat_optics) is pure Python and has no FFI; the
 * boundary it exposes is the Python function surface its notebooks call.  Each entry point
 * below names the reference function (file:line, relative to the reference checkout) whose
 * arithmetic it replaces.  sosat_optics_b200/{ray_trace,opt_analyze}.py bind these with
 * ctypes under the reference's own function names; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 *   - every function returns 0 on success or a negative sosat_status_t; the message is
 *     available from sosat_last_error() (thread-local).
 *   - functions taking a stream are asynchronous on it (stream = cudaStream_t as void*;
 *     NULL = legacy default stream).  The *_host convenience calls are synchronous.
 *   - the caller owns every buffer.  The library keeps only the per-device geometry in
 *     __constant__ memory (sosat_set_geometry) and allocates nothing persistent; the
 *     *_host calls allocate and free their own temporaries.
 *   - numerical invalidity is not an error: an invalid ray gets amplitude = OPL = 0 exactly
 *     like ray_trace.py:1578-1580.
 *   - there is no CPU fallback anywhere in this library.
 *   - threading: one host process per GPU is the intended use (torch.distributed for several).
 *     Calls on one device may come from several host threads; the plan registry of the DFT
 *     workspaces is mutex-protected.  The geometry, the per-call frequency table and the strip
 *     queue of the fused trace + bin kernel are per-DEVICE state written in stream order: trace
 *     calls must not overlap on different streams of the same device (one after the other on
 *     any stream is fine).
 *     A DFT workspace (scratch of the transform in flight + the twiddles cached in it, written
 *     by a kernel on the stream of the first call) belongs to ONE stream: give every stream its
 *     own workspace (the Python wrappers key their workspace cache by device and stream).
 */
#ifndef SOSAT_B200_H
#define SOSAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOSAT_ABI_VERSION 1
#define SOSAT_NUM_SURFACES 6
#define SOSAT_OUT_ROWS 17 /* ray_trace.py:1095  out = np.zeros((17, n_pts)) */
#define SOSAT_MAX_FREQS 64

typedef enum {
    SOSAT_OK = 0,
    SOSAT_ERR_INVALID_ARG = -1,
    SOSAT_ERR_CUDA = -2,
    SOSAT_ERR_NO_GEOMETRY = -3,
    SOSAT_ERR_UNSUPPORTED = -4,
    SOSAT_ERR_NO_DEVICE = -5
} sosat_status_t;

/* One refracting surface, in trace order 3b,3a,2b,2a,1b,1a.
 * sag_mm(x,y) = 10*[c r^2/(1+sqrt(1-(1+k)c^2 r^2)) + a1 r^2 + a2 r^4 + a3 r^6], r in cm
 * (ot_geo.py:49-91,182-227,313-354; slopes :94-143,230-276,364-414).  The surface frame is
 * the tele frame translated to y0 and rotated by pi/2 about x (ot_geo.py:162-179 etc.). */
typedef struct {
    double c, k, a1, a2, a3;
    double y0;          /* vertex position on the tele y axis [mm] (ot_geo.py:43-46)        */
    double n_in, n_out; /* refractive index before / after the surface (ot_geo.py:13-14)    */
    double t_lo, t_hi;  /* the reference's brentq bracket [mm] (ray_trace.py:1130,1198,...) */
} sosat_surface_t;

/* Frequency-independent optics (SatGeo + ot_geo module constants). */
typedef struct {
    sosat_surface_t surf[SOSAT_NUM_SURFACES];
    double y_lyot;    /* ot_geo.py:46                                              */
    double y_source;  /* SatGeo.y_source (ot_geo.py:29)                            */
    double clip_l3;   /* 392/2 mm: lens-3 clip, strict >= (ray_trace.py:1138)      */
    double clip_stop; /* 210 mm: Lyot-stop and lens-2a clip, <= (ray_trace.py:1569) */
    double cone;      /* 0.4 rad launch half-cone (ray_trace.py:1086-1087)         */
} sosat_geom_t;

/* Per-frequency feed/phase parameters (SatGeo.th_fwhp_x, th_fwhp_y, k). */
typedef struct {
    double fwhp_x;    /* rad; divides de_ho (ray_trace.py:1575) */
    double fwhp_y;    /* rad; divides de_ve                      */
    double k_over_2e3;/* SatGeo.k / 1e3 / 2: phase per mm of OPL (ray_trace.py:1601) */
} sosat_freq_t;

/* Diagnostics accumulated by the tracer (device array of SOSAT_NUM_STATS doubles). */
enum {
    SOSAT_STAT_N_VALID = 0,  /* rays with amplitude != 0                          */
    SOSAT_STAT_SUM_OPL = 1,  /* sum of OPL over valid rays [mm]                    */
    SOSAT_STAT_SUM_DX = 2,   /* sum of exit direction x,y,z over valid rays        */
    SOSAT_STAT_SUM_DY = 3,
    SOSAT_STAT_SUM_DZ = 4,
    SOSAT_STAT_N_SLOW = 5,   /* rays re-solved with the bracketed (Brent) slow path */
    SOSAT_STAT_N_BRACKET_FAIL = 6, /* solves where the reference's brentq would raise */
    SOSAT_STAT_N_BINNED = 7, /* valid rays that landed inside the bin grid          */
    SOSAT_NUM_STATS = 8
};

/* ---- library / device ------------------------------------------------------------- */
int sosat_abi_version(void);
/* Bit 0 (SOSAT_BUILD_EXPERIMENTS): the library was compiled with -DSOSAT_EXPERIMENTS and also
 * holds the kernels that lost their A/B runs (other Newton schedules, split-K / one-tile-per-CTA
 * far-field kernels, DESIGN.md sections 3-4), selectable by SOSAT_TRACE_VARIANT / SOSAT_GEMM_*
 * environment variables.  The product library is built without them. */
#define SOSAT_BUILD_EXPERIMENTS 1
int sosat_build_flags(void);
const char *sosat_last_error(void);
int sosat_device_count(void);
int sosat_set_device(int device);
/* name: >= 256 bytes; sm: compute capability major*10+minor; n_sm: multiprocessors */
int sosat_device_info(char *name, int *sm, int *n_sm, size_t *mem_bytes);

/* ---- geometry (ot_geo.py:8-46) ------------------------------------------------------ */
/* Copies *h_geom into the current device's __constant__ memory (async on stream). */
int sosat_set_geometry(const sosat_geom_t *h_geom, void *stream);

/* ---- stage 1: ray trace (ray_trace.py:1070-1591 rx_to_lyot) ------------------------- */
/* Rays ii in [ray_begin, ray_end) of the n_scan x n_scan launch grid, ii = iphi*n_scan +
 * itheta (ray_trace.py:1086-1091), from feed position h_rx[3] (tele frame, mm).
 * d_out: [17][ray_end-ray_begin] row-major float64, rows as ray_trace.py:1558-1588
 * (0-2 pos_ap, 3 OPL, 4 amplitude, 5-7 last normal, 8-10 exit direction, 11-16 zero).
 * d_stats: SOSAT_NUM_STATS doubles, accumulated into (caller zeroes), may be NULL. */
int sosat_trace_rays(const double *h_rx, const sosat_freq_t *h_freq, int n_scan,
                     int64_t ray_begin, int64_t ray_end, double *d_out, double *d_stats,
                     void *stream);

/* Compact per-ray output of the same trace (SURVEY 8d "compact SoA mode", 16 B/ray): for every
 * ray one float4 {x_sim [cm], y_sim [cm], a cos p, a sin p}, p = k_over_2e3 (OPL - opl_ref),
 * all-zero for invalid rays -- the sample format sosat_nudft_direct consumes, so a near field on
 * the scattered ray positions goes to the far field without gridding.  d_xyb: [n][4] float32,
 * 16-byte aligned. */
int sosat_trace_samples(const double *h_rx, const sosat_freq_t *h_freq, int n_scan,
                        int64_t ray_begin, int64_t ray_end, double opl_ref, float *d_xyb,
                        double *d_stats, void *stream);

/* Host-buffer form of the same call: h_out is [17][n] float64 host memory.  Synchronous. */
int sosat_rx_to_lyot_host(const sosat_geom_t *h_geom, const double *h_rx,
                          const sosat_freq_t *h_freq, int n_scan, int64_t ray_begin,
                          int64_t ray_end, double *h_out, double *h_stats);

/* ---- stage 1b: getNearField post-processing (ray_trace.py:1594-1614) ---------------- */
/* From d_out (layout above, n rays) computes, for every ray, x_sim = out[0]/10,
 * y_sim = out[2]/10, a_sim = out[4], p_sim = mod(k (OPL - <OPL>)/1e3/2, 2 pi) - <p>, where
 * the means run over rays with out[4] != 0, and d_mask[i] = (out[4][i] != 0).
 * d_xyap: [4][n] float64 (x, y, a, p); d_tan_og: 3 doubles (mean exit direction);
 * d_work: SOSAT_NUM_STATS doubles of scratch.  Compaction by d_mask is the caller's. */
int sosat_near_field_arrays(const double *d_out, int64_t n, double k, double *d_xyap,
                            uint8_t *d_mask, double *d_tan_og, double *d_work,
                            void *stream);

/* ---- stage 1c: fused trace + near-field binning (north_star) -- */
/* Traces launch rows iphi in [row_begin, row_end) (all itheta) for n_rx feed positions
 * (d_rx: [n_rx][3] device doubles) and n_freq <= SOSAT_MAX_FREQS frequencies, and
 * accumulates, per (rx, freq), sum a*cos(p), sum a*sin(p) and the ray count into
 * grid_n x grid_n float32 planes covering [-extent_mm/2, extent_mm/2)^2 of the source
 * plane: row index = tele z (reference y_sim), column index = tele x (x_sim).
 * p = k_over_2e3 * (OPL - opl_ref).  Planes are accumulated into (caller zeroes):
 *   d_re, d_im: [n_rx][n_freq][grid_n][grid_n];  d_cnt: [n_rx][grid_n][grid_n];
 *   d_stats:    [n_rx][SOSAT_NUM_STATS] doubles.
 * d_re/d_im/d_cnt may all be NULL to accumulate the statistics only. */
int sosat_trace_bin(const double *d_rx, int n_rx, const sosat_freq_t *h_freq, int n_freq,
                    int n_scan, int64_t row_begin, int64_t row_end, int grid_n,
                    double extent_mm, double opl_ref, float *d_re, float *d_im, float *d_cnt,
                    double *d_stats, void *stream);

/* The same for an arbitrary set of TILE ROWS of the launch grid: tile row t = launch rows
 * [SOSAT_TRACE_TILE_ROWS t, SOSAT_TRACE_TILE_ROWS (t + 1)) (the last one may be ragged);
 * d_tile_rows: n_tile_rows DEVICE ints, each in [0, ceil(n_scan / SOSAT_TRACE_TILE_ROWS)), no
 * duplicates (a duplicate is traced twice).  Multi-GPU ray sharding deals the tile rows cyclically
 * to the ranks so that every rank traces the same mix of central and marginal rays. */
#define SOSAT_TRACE_TILE_ROWS 16
int sosat_trace_bin_tiles(const double *d_rx, int n_rx, const sosat_freq_t *h_freq, int n_freq,
                          int n_scan, const int *d_tile_rows, int n_tile_rows, int grid_n,
                          double extent_mm, double opl_ref, float *d_re, float *d_im, float *d_cnt,
                          double *d_stats, void *stream);

/* b = (re + i im)/cnt * exp(-i k_over_2e3 (mean_opl - opl_ref)) per cell (0 where cnt = 0);
 * d_b: [n][2] float32 interleaved complex.  In place over (re, im) is not supported. */
int sosat_bins_to_field(const float *d_re, const float *d_im, const float *d_cnt, int64_t n,
                        double phase_offset_rad, float *d_b, void *stream);

/* The same with the phase reference taken on the device: phase_offset = k_over_2e3 *
 * (mean OPL - opl_ref), mean OPL = d_stats[SUM_OPL] / d_stats[N_VALID] of the (already reduced)
 * statistics row of this receiver -- no host round trip between the trace and the far field. */
int sosat_bins_to_field_stats(const float *d_re, const float *d_im, const float *d_cnt, int64_t n,
                              const double *d_stats, double k_over_2e3, double opl_ref,
                              float *d_b, void *stream);
/* The whole stack of fields of a sweep in ONE launch (grid y = field): d_re / d_im / d_b
 * [n_rx][n_freq][n], d_cnt [n_rx][n], d_stats [n_rx][SOSAT_NUM_STATS], h_k_over_2e3 [n_freq]
 * (HOST array, copied into the launch parameters).  Replaces the n_rx x n_freq Python loop over
 * sosat_bins_to_field_stats (a 32-receiver x 64-frequency sweep is 2 048 fields). */
int sosat_bins_to_field_stats_batch(const float *d_re, const float *d_im, const float *d_cnt,
                                    int64_t n, int n_rx, int n_freq, const double *d_stats,
                                    const double *h_k_over_2e3, double opl_ref, float *d_b,
                                    void *stream);

/* ---- stage 2: far-field transform (opt_analyze.py:220-230 a2b) ----------------------- */
/* B = W_r b W_c^T, W[k,n] = exp(-2 pi i (k - floor(N/2)) (n - ceil(N/2)) / N): identical to
 * fftshift(fft2(fftshift(b))).  Dense separable complex contraction executed as two
 * split-bf16 GEMM passes on tcgen05 tensor cores with TMA-staged tiles.
 * d_b, d_B: [n][n][2] float32 interleaved complex (may alias).  Workspace from
 * sosat_dft2_workspace_bytes(n), 256-byte aligned; twiddle operands are cached in it. */
size_t sosat_dft2_workspace_bytes(int n);
/* The twiddle operands written into a workspace are remembered per (device, address, n);
 * call this before freeing a workspace so a later allocation at the same address is
 * re-planned. */
int sosat_dft2_forget(const void *d_workspace);
int sosat_dft2_gemm(const float *d_b, int n, float *d_B, void *d_workspace,
                    size_t workspace_bytes, void *stream);

/* The other centred transforms of opt_analyze.py on the same GEMM path (a workspace holds the
 * twiddles of one kind at a time and is re-planned when the kind changes):
 *   SOSAT_DFT_FORWARD  fftshift(fft2(fftshift(x)))      a2b :225-228, beam_convolve :76-82
 *   SOSAT_DFT_INVERSE  ifftshift(ifft2(ifftshift(x)))   beam_convolve :86-88
 *   SOSAT_DFT_B2A      fftshift(ifft2(x))               b2a :213-215 (line :213's shifted copy
 *                                                       is discarded by :214; quirk kept) */
enum { SOSAT_DFT_FORWARD = 0, SOSAT_DFT_INVERSE = 1, SOSAT_DFT_B2A = 2 };
int sosat_dft2_gemm_kind(const float *d_b, int n, float *d_B, int kind, void *d_workspace,
                         size_t workspace_bytes, void *stream);

/* Batched form for frequency / receiver sweeps: `batch` fields of one size [batch][n][n] through
 * one transpose-and-split kernel and ONE launch per UMMA pass (the batch supplies the tiles that
 * a single 1024^2 transform lacks to fill the machine with cta_group::2 pair tiles).  d_B:
 * [batch][n][n][2] float32.  Workspace from sosat_dft2_batch_workspace_bytes(n, batch). */
size_t sosat_dft2_batch_workspace_bytes(int n, int batch);
int sosat_dft2_gemm_batch(const float *d_b, int n, int batch, float *d_B, int kind,
                          void *d_workspace, size_t workspace_bytes, void *stream);
/* The same for ZERO-PADDED fields (zero_pad, opt_analyze.py:20-32, before a2b; beam_sweep's pad):
 * d_b [batch][n_in][n_in] is transformed as if it sat at rows / columns [pad, pad + n_in) of an
 * n x n array of zeros, n = n_in + 2 pad; d_B [batch][n][n].  The zeros are not multiplied: both
 * passes contract over the n_in columns of W that meet data (3/8 of the work at n = 2 n_in) and
 * the padded input is never materialised.  Needs n padded to 128 to be a multiple of 256. */
size_t sosat_dft2_batch_padded_workspace_bytes(int n_in, int pad, int batch);
int sosat_dft2_gemm_batch_padded(const float *d_b, int n_in, int pad, int batch, float *d_B, int kind,
                                 void *d_workspace, size_t workspace_bytes, void *stream);

/* Zoomed far field on caller-chosen angles (the matrix-DFT form's advantage over
 * an FFT): for j, k < n_out
 *   B[j][k] = sum_{m,n} b[m][n] exp(-2 pi i (y_m thy_j + x_n thx_k) * inv_lambda)
 * with row coordinates h_y[n], column coordinates h_x[n] (any spacing, unit of 1/inv_lambda) and
 * angles h_thy[n_out], h_thx[n_out] [rad] (host arrays, n_out <= n rounded up to 128).  Same two
 * UMMA passes as sosat_dft2_gemm with twiddles generated for these angles on every call.
 * d_B: [n_out][n_out][2] float32.  With x_n = (n - ceil(N/2)) dx, th_k = (k - floor(N/2))
 * lambda / (N dx) it equals sosat_dft2_gemm. */
int sosat_dft2_zoom(const float *d_b, int n, float *d_B, int n_out, const double *h_y,
                    const double *h_x, const double *h_thy, const double *h_thx,
                    double inv_lambda, void *d_workspace, size_t workspace_bytes, void *stream);

/* a2b with the reference's argument meaning: d_apert (amplitude, abs() is taken) and
 * d_phi_deg (phase in degrees) are [n][n] float64; outputs d_beam [n][n][2] float64
 * (complex128) and d_phase [n][n] float64 = arctan2(imag, real). */
int sosat_a2b(const double *d_apert, const double *d_phi_deg, int n, double *d_beam,
              double *d_phase, void *d_workspace, size_t workspace_bytes, void *stream);
/* b2a (opt_analyze.py:208-217), same argument meaning as the reference: |beam| and phase in
 * degrees in, complex128 aperture field and its arctan2 phase out. */
int sosat_b2a(const double *d_beam, const double *d_phase_deg, int n, double *d_aper,
              double *d_aper_phase, void *d_workspace, size_t workspace_bytes, void *stream);
/* beam_convolve building blocks (opt_analyze.py:60-88): the cos-tapered rectangular horn
 * aperture on the caller's coordinate grids (complex64 out, n = number of cells) and a
 * pointwise complex product; the three transforms go through sosat_dft2_gemm_kind. */
int sosat_horn_aperture(const double *d_x, const double *d_y, int64_t n, double apert1,
                        double apert2, float *d_out, void *stream);
int sosat_cmul(const float *d_a, const float *d_b, int64_t n, float *d_out, void *stream);
/* Host-buffer form (synchronous): allocates and frees device temporaries. */
int sosat_a2b_host(const double *h_apert, const double *h_phi_deg, int n, double *h_beam,
                   double *h_phase);

/* Direct-sum far field on irregular samples (north_star): for j < n_ang
 *   B_j = sum_p b_p exp(-2 pi i (x_p thx_j + y_p thy_j) * inv_lambda),
 * phase factor generated in registers.  d_xyb: [n_pts] float4 {x, y, Re b, Im b} (x, y in
 * the unit of 1/inv_lambda); d_th: [n_ang] float2 {thx, thy} [rad]; d_B: [n_ang] float2. */
int sosat_nudft_direct(const float *d_xyb, int64_t n_pts, const float *d_th, int64_t n_ang,
                       double inv_lambda, float *d_B, void *stream);

/* ---- on-device reductions for sweeps -------------------------- */
/* get_beam_cent / get_fwhm / tan_og of getNearField (ray_trace.py:1604-1614, 1652-1666) from the
 * ray table d_out ([17][n], rays ray_begin .. ray_begin+n of the n_scan^2 launch grid), for
 * n_freq feed patterns at once (the ray geometry is frequency independent; the amplitude of
 * ray (itheta, iphi) follows from its launch angles).  d_result (6 + 2*SOSAT_MAX_FREQS doubles):
 *   [0] n_valid  [1] mean x_sim [cm]  [2] mean y_sim [cm]  [3..5] tan_og
 *   [6+2f] fwhm_nf_x(f) = max|y_sim - cy|, [7+2f] fwhm_nf_y(f) = max|x_sim - cx| over rays with
 *   a_f^2 > max(a_f^2)/2 (the reference's swapped naming kept).
 * d_work: 8 + SOSAT_MAX_FREQS doubles of scratch. */
int sosat_near_field_metrics(const double *d_out, int64_t n, int n_scan, int64_t ray_begin,
                             double cone, const sosat_freq_t *h_freq, int n_freq,
                             double *d_result, double *d_work, void *stream);
/* The searches of ff_fwhm_est (ray_trace.py:1757-1779) on a complex64 far field d_B [n][n]:
 * d_idx[6] = {peak row, peak column (first maximum of |B|^2, row-major), first / last column
 * above half power in the peak row, first / last row above half power in the peak column};
 * d_val[0] = peak |B|^2.  d_work: 8 bytes of scratch. */
int sosat_ff_fwhm_indices(const float *d_B, int n, int *d_idx, float *d_val, void *d_work,
                          void *stream);
/* The same for a stack d_B [n_fields][n][n] in one launch per search: d_idx [n_fields][6],
 * d_val [n_fields], d_work 8 * n_fields bytes.  A field without a finite non-zero cell (e.g. a
 * NaN in the near field, which the contraction spreads over the whole far field) reports all six
 * indices as -1 and val = 0; the Python wrappers raise on it. */
int sosat_ff_fwhm_indices_batch(const float *d_B, int n, int n_fields, int *d_idx, float *d_val,
                                void *d_work, void *stream);

/* ---- scattered rays -> regular grid -> stage grid --------------------------
 * ray_trace.py:1675-1717 sim2d: scipy.interpolate.griddata(method="cubic") (Clough-Tocher on the
 * Delaunay triangulation, globally estimated gradients) of amplitude and phase of the rays within
 * r_select of the beam centre; ray_trace.py:1720-1754 sim_data: interp2d(kind="cubic") resampling.
 * SciPy is not vendored by the reference (unpinned); csrc/gridder.cu restates the published
 * algorithms for rays that form a warped structured lattice (d_ray_index[i] = iphi * n_scan +
 * itheta of point i), see the header of that file.
 * sosat_sim2d_prepare: beam centre, selection, lattice triangulation, `sweeps` four-colour
 *   Gauss-Seidel sweeps of the gradient estimate (scipy converges to 1e-6 in 10-13).  d_x, d_y,
 *   d_a, d_p [n_pts] doubles (x_sim, y_sim [cm], a_sim, p_sim).  d_scalars[8] receives {cx, cy,
 *   xmin, xmax, ymin, ymax of the selected points, their number, lattice index of the centre ray}.
 * sosat_sim2d_eval: the interpolant at the m x m points (d_gx[i], d_gy[j]) -> d_out_a / d_out_p
 *   [m][m] (x along axis 0, like np.mgrid); 0 outside the triangulation and for r >= r_zero. */
size_t sosat_sim2d_workspace_bytes(int64_t n_pts, int n_scan);
int sosat_sim2d_prepare(const double *d_x, const double *d_y, const double *d_a, const double *d_p,
                        const int64_t *d_ray_index, int64_t n_pts, int n_scan, double r_select,
                        int sweeps, void *d_workspace, size_t workspace_bytes, double *d_scalars,
                        void *stream);
int sosat_sim2d_eval(const double *d_x, const double *d_y, const double *d_a, const double *d_p,
                     int64_t n_pts, int n_scan, const void *d_workspace, const double *d_gx,
                     const double *d_gy, int m, double r_zero, double *d_out_a, double *d_out_p,
                     void *stream);
/* sim_data: z_amp = |a|^T / max|a|, z_ph = arg((a e^{ip})^T) of the n x n grid, resampled as
 * My z Mx^T with the m x n spline operators of the two axes (built by the caller: FITPACK's
 * not-a-knot interpolating cubic spline is linear in the data), a_data = resampled amplitude /
 * its maximum, p_data = resampled phase.  d_work: sosat_sim_data_work_doubles(n, m) doubles. */
size_t sosat_sim_data_work_doubles(int n, int m);
int sosat_sim_data(const double *d_a, const double *d_p, int n, const double *d_Mx,
                   const double *d_My, int m, double *d_a_data, double *d_p_data, double *d_work,
                   void *stream);

/* ---- measurement helpers --------------------------------------------------------------- */
/* ---- multi-GPU: reduce of the bin planes over NVLink peer memory -----------
 * The reference has no multi-device path; north_star shards the rays of a bundle over the GPUs
 * of a node and reduces the bins once.  sosat_peer_* map one rank's accumulator buffer into the
 * other ranks' address spaces (CUDA IPC; one process per GPU); sosat_bins_reduce_field is the
 * reduce fused with its consumer (sosat_bins_to_field_stats' normalisation, the replacement of
 * getNearField's post-processing, ray_trace.py:1599-1614) and with the gather of the result. */
#define SOSAT_PEER_HANDLE_BYTES 64
/* Exportable device buffer (its own cudaMalloc, zero-filled) / release. */
int sosat_peer_alloc(int64_t bytes, void **d_ptr);
int sosat_peer_free(void *d_ptr);
/* handle: SOSAT_PEER_HANDLE_BYTES bytes to send to the other processes of the node. */
int sosat_peer_export(const void *d_ptr, unsigned char *handle);
/* Map a peer's buffer (handle from ANOTHER process); peer access is enabled on first use. */
int sosat_peer_open(const unsigned char *handle, void **d_peer);
int sosat_peer_close(void *d_peer);
/* For the cells [cell_begin, cell_end) of one (receiver, frequency) field: sum re / im / count
 * over the n_peers buffers (float32 planes at byte offsets re_off / im_off / cnt_off from each
 * base, identical layout everywhere), sum their statistics rows (SOSAT_NUM_STATS doubles at
 * stats_off), normalise as sosat_bins_to_field_stats does with the SUMMED statistics, and store
 * the complex64 cells at byte offset out_off of each of the n_out destination buffers.
 * h_peer_base / h_out_base: HOST arrays of device pointers (own buffer and sosat_peer_open
 * mappings).  d_stats_sum (optional, local): receives the summed statistics row.  The caller
 * orders the launch after every peer's trace and the next overwrite of the planes after every
 * peer's launch (stream-ordered barriers).  At most 16 peers / destinations. */
int sosat_bins_reduce_field(const void *const *h_peer_base, int n_peers, int64_t re_off,
                            int64_t im_off, int64_t cnt_off, int64_t stats_off,
                            int64_t cell_begin, int64_t cell_end, double k_over_2e3,
                            double opl_ref, void *const *h_out_base, int n_out, int64_t out_off,
                            double *d_stats_sum, void *stream);
/* The same for the whole stack of (receiver, frequency) fields of a sweep in one launch (grid
 * y = field irx * n_freq + f): planes of n_cells cells each, re / im / field stacked
 * [n_rx][n_freq], counts and statistics rows [n_rx]; offsets are those of field 0;
 * h_k_over_2e3 [n_freq] HOST array; d_stats_sum [n_rx][SOSAT_NUM_STATS].  With more than one
 * field n_cells must be a multiple of 4 (16-byte aligned planes). */
int sosat_bins_reduce_fields(const void *const *h_peer_base, int n_peers, int64_t re_off,
                             int64_t im_off, int64_t cnt_off, int64_t stats_off, int64_t n_cells,
                             int n_rx, int n_freq, int64_t cell_begin, int64_t cell_end,
                             const double *h_k_over_2e3, double opl_ref,
                             void *const *h_out_base, int n_out, int64_t out_off,
                             double *d_stats_sum, void *stream);

/* Stream-ordered barrier over the ranks of the node through flags in the peer-mapped buffers:
 * 64 uint64 slots at byte offset flags_off of every buffer (zero-initialised by sosat_peer_alloc;
 * slot 63 records a timeout).  Every rank calls it with the same, ever-increasing epoch (1, 2, ...).
 * Replaces a one-element NCCL all-reduce: 8 us instead of 40 us at 8 GPUs.  A rank that waits
 * longer than timeout_s (<= 0: 10 s) gives up and stores the epoch in slot 63 instead of hanging. */
int sosat_peer_barrier(const void *const *h_peer_base, int n_peers, int rank, int64_t flags_off,
                       uint64_t epoch, double timeout_s, void *stream);

/* Dependent-free FP64 FMA throughput microkernel: returns achieved FLOP/s (2 per FMA) in
 * *flops_per_s; used by bench.py as the roofline denominator of the ray tracer because
 * MEASURED_PEAKS.json carries no FP64 peak. */
int sosat_measure_fp64_peak(double *flops_per_s, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* SOSAT_B200_H */
