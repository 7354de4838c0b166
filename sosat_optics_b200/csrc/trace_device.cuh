// Device-side ray tracer: one ray per thread, fp64 state, geometry in __constant__ memory.
//
// Replaces the per-ray body of ray_trace.py:1098-1588 (rx_to_lyot).  The six surface
// intersections the reference finds with scipy.optimize.brentq (ray_trace.py:1130, 1198,
// 1261, 1326, 1393, 1460) are found by Newton iteration on the same function
//     f(t) = z_l(o + d t) - sag(x_l, y_l)              (root_z?? closures, e.g. :1114-1128)
// written directly in the tele frame: with every tilt angle = pi/2 (ot_geo.py:32-37) the
// surface frame is x_l = x, y_l = z, z_l = -(y - y0) (ot_geo.py:449-458), so
//     f(t) = (y0 - o_y) - d_y t - S(u),   u = X^2 + Z^2 [mm^2],  X = o_x + d_x t, Z = o_z + d_z t
//     S(u) = u [ c'/(1+s) + a1' + a2' u + a3' u^2 ],  s = sqrt(1 - K' u)
// with the cm->mm factors of ot_geo.py:54-55,69 folded into the primed constants.  Two exact
// identities remove the reciprocal from the evaluation (1 - s^2 = K'u):
//     u c'/(1+s) = (c'/K') (1 - s)              conic sag without a division
//     dS/du * 2  = c'/s + 2 a1' + 4 a2' u + 6 a3' u^2    (the textbook c r / sqrt(1 - K r^2))
// The first needs |c'/K'| = 10/|(1+k) c| mm to stay moderate (the rounding of s is amplified by
// it); a near-parabolic surface (|c'/K'| > 1e6 mm, or K' = 0) selects the GENERIC kernels, which
// keep the u c'/(1+s) form.
// A ray whose Newton solve is not clearly converged, or that grazes the conic-domain edge
// of lens 1b, is re-traced with a bracketed Brent solve that restates the reference's
// brentq call (same brackets, same tolerances), so such rays get the reference's answer.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "sosat_b200.h"

namespace sosat {

struct SurfC {
    double Kp;    // 0.01 (1+k) c^2
    double cp;    // 0.1 c
    double a1p;   // 0.1 a1
    double a2p;   // 1e-3 a2
    double a3p;   // 1e-5 a3
    double a1p2;  // 2 a1'
    double a2p4;  // 4 a2'
    double a3p6;  // 6 a3'
    double y0;    // vertex y [mm]
    double mu;    // n_in / n_out
    double mu2;   // mu^2
    double n_in;  // OPL weight of the segment arriving at this surface
    double t_lo, t_hi;
    double cK;    // c'/K' = 10 / ((1+k) c) [mm]: conic sag = cK (1 - s); 0 when K' = 0
    double omm2;  // 1 - mu^2
    int refine;   // |cK| > 300 mm: s needs the fully refined rsqrt in the division-free form
    int pad_;
    double cKh;    // c' K' / 2: d(c'/s)/du = cKh / s^3
    double y0K;    // y0 - cK
};

struct SurfF {
    float Kp, cp, a1p, a2p, a3p, a1p2, a2p4, a3p6, cK, cKh;
};

struct GeomC {
    SurfC s[SOSAT_NUM_SURFACES];
    SurfF f[SOSAT_NUM_SURFACES];
    double y_lyot, y_source, clip_l3_sq, clip_stop_sq, cone;
};

// Newton acceptance.  The surface normal is taken at the iterate before the last step, so the
// last step bounds the normal error (curvature ~1e-2/mm, up to ~1e2/mm for wild rays near the
// lens-1b conic edge).  After the fixed schedule a ray whose last step exceeds REFINE is
// re-traced by the adaptive path (up to SOSAT_NEWTON_EXTRA more fp64 steps per surface, rare,
// lane-divergent); one still above TOL (or NaN) is marked for the bracketed slow path.
#define SOSAT_NEWTON_REFINE 1.0e-9
#define SOSAT_NEWTON_TOL 1.0e-8
#define SOSAT_NEWTON_EXTRA 6
// lens 1b is the only surface with a bounded conic domain (k > -1): hits
// with w = 1 - K'u below this are within ~8 mm of the domain edge r = 300.4 mm, where the
// reference's brentq converges onto the sag discontinuity instead of a root.
#define SOSAT_EDGE_W 0.05
// Schedule "N32 + 1": ONE fp64 Newton step from the fp32 root.  Its error is kappa step^2 with
// kappa = |f''/2f'| ~ 1e-3/mm on these surfaces, so a step below this bound leaves <= 1e-9 mm;
// longer steps send the lane to the adaptive path.  The fp32 root sits at the fp32 floor
// (steps of 1e-5 .. 3e-4 mm measured), so the bound is not on the hot path of any ray.
// The bound is a power of two so that it can be tested on the exponent field: 2^-10 mm.
#ifndef SOSAT_LINEAR_MAX_STEP_LOG2
#define SOSAT_LINEAR_MAX_STEP_LOG2 10
#endif

enum : unsigned {
    RAY_CLIPPED_L3 = 1u,  // ray_trace.py:1138-1140: whole output column stays zero
    RAY_FLAGGED = 2u,     // Newton not clearly converged / conic edge: needs the slow path
    RAY_BRACKET_FAIL = 4u,  // slow path: f(lo), f(hi) same sign (reference raises ValueError)
    RAY_UNCONVERGED = 8u    // fast path: last fixed Newton step above REFINE, needs the adaptive path
};

struct RayResult {
    double ax, ay, az;  // pos_ap (source plane), mm
    double opl;         // total optical path length, mm
    double nx, ny, nz;  // last surface normal, tele frame (out rows 5-7)
    double dx, dy, dz;  // exit direction (out rows 8-10)
    unsigned flags;
    bool valid;         // passes the stop / lens-2a clips, nothing NaN
};

// ---- fast fp64 reciprocal / reciprocal square root: MUFU seed + Newton-Raphson ---------
__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
// Measured on B200 (parity runs in DESIGN.md section 3): the rsqrt seed is good to ~1e-6 and
// the rcp seed only to ~3e-5.  rsqrt_fast = seed + one Newton-Raphson step (~3e-12 relative:
// <= 2e-10 mm on a 50 mm sag, far inside the 1e-3 rad phase tolerance) at half the DFMA count
// of a fully rounded result; rcp_full = seed + two steps (~1e-18); the raw rcp seed is used
// only for the Newton step quotient, where a 3e-5 relative error merely makes the last
// iterations converge linearly with ratio 3e-5 instead of quadratically.
__device__ __forceinline__ double rcp_full(double x) {
    double r = rcp_seed(x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double rcp_half(double x) {
    double r = rcp_seed(x);
    double e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y = rsqrt_seed(x);
    double a = x * y;
    double e = fma(-a, y, 1.0);
    return fma(0.5 * y, e, y);
}
// Full precision in ONE cubically convergent step from the seed (5 DP instead of the 8 of two
// Newton-Raphson steps): with e = 1 - x y^2 (|e| ~ 2e-6), 1/sqrt(x) = y (1 + e/2 + 3 e^2/8 + O(e^3)),
// remainder 5 e^3/16 ~ 3e-18.
__device__ __forceinline__ double rsqrt_full(double x) {
    const double y = rsqrt_seed(x);
    const double e = fma(-(x * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
}
// Where each precision is used (measured against the oracle on 2e6 rays):
// the sag evaluation tolerates the one-step rsqrt (median |dOPL| 2e-12 mm), the normalisation
// and the Snell radicand do not (one-step there: median 8e-9 mm, 3e-4 mm on wild marginal rays).
#ifndef SOSAT_RSQRT_EVAL
#define SOSAT_RSQRT_EVAL rsqrt_fast
#endif
#ifndef SOSAT_REFRACT_JOINT
#define SOSAT_REFRACT_JOINT 1
#endif
#ifndef SOSAT_LINEAR_RCP
#define SOSAT_LINEAR_RCP rcp_seed
#endif
#ifndef SOSAT_RSQRT_SNELL
#define SOSAT_RSQRT_SNELL rsqrt_full
#endif
// double <-> float for the SEEDS of the fp32 Newton steps.  F2F.F32.F64 costs ~9 cycles of the XU
// pipe per warp instruction on B200 (scripts/xu_mix.cu) -- the pipe the MUFUs of the fp32
// pre-iterations need -- so an integer bit-manipulation form (truncating, tiny values flushed
// towards 0; NaN / Inf give garbage, harmless because the fp64 state they came from stays
// NaN / Inf) is compiled with -DSOSAT_FAST_CVT=1.  It is parity-green but its 5-6 ALU
// instructions per conversion cost more issue slots than the XU time they free: 1.67 -> 1.55e10
// rays/s.  Off by default.
#ifndef SOSAT_FAST_CVT
#define SOSAT_FAST_CVT 0
#endif
__device__ __forceinline__ float d2f_seed(double x) {
#if SOSAT_FAST_CVT
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = max((hi & 0x7fffffff) - 0x38000000, 0);  // rebias the exponent (1023 -> 127)
    const unsigned f = __funnelshift_l((unsigned)lo, (unsigned)e, 3) | ((unsigned)hi & 0x80000000u);
    return __uint_as_float(f);
#else
    return (float)x;
#endif
}
__device__ __forceinline__ double f2d_seed(float x) {
#if SOSAT_FAST_CVT
    const unsigned f = __float_as_uint(x), m = f & 0x7fffffffu;
    const unsigned hi = (f & 0x80000000u) | ((m >> 3) + (m ? 0x38000000u : 0u));
    return __hiloint2double((int)hi, (int)(m << 29));
#else
    return (double)x;
#endif
}
__device__ __forceinline__ float rcp_f32(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_f32(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- surface function in the tele frame -------------------------------------------------
// Returns f and df/dt at parameter t, plus X, Z of the point and G = (dS/dX)/X = (dS/dZ)/Z,
// the slope factor of d_z?? (ot_geo.py:108-115): the surface gradient is (G X, G Z).
// y0PyK = y0Py - cK in the division-free form (GENERIC: y0PyK = y0Py): the conic term is then
// f = [y0PyK - d_y t - u poly] + (cK w) rs, i.e. one DFMA after the rsqrt chain ends, with the
// bracket and cK w evaluated while the MUFU seed is in flight.
template <bool GENERIC>
__device__ __forceinline__ void surf_eval(const SurfC &S, const bool refine, double Px, double Pz, double dx,
                                          double dy, double dz, double y0PyK, double t,
                                          double &f, double &fp, double &X, double &Z,
                                          double &G, double &w, double &u, double &rs,
                                          double &xd, double &h) {
    X = fma(dx, t, Px);
    Z = fma(dz, t, Pz);
    u = fma(X, X, Z * Z);
    w = fma(-S.Kp, u, 1.0);
    // |cK| > 300 mm (refine folds to a constant in the unrolled surface loop): the (1 - s)
    // cancellation needs the fully rounded rsqrt there
    rs = (!GENERIC && refine) ? rsqrt_full(w) : SOSAT_RSQRT_EVAL(w);
    const double poly = fma(fma(S.a3p, u, S.a2p), u, S.a1p);
    const double base = fma(-u, poly, fma(-dy, t, y0PyK));
    if (GENERIC) {
        const double s = w * rs;
        f = fma(-u, S.cp * rcp_full(1.0 + s), base);
    } else {
        f = fma(S.cK * w, rs, base);  // ... - cK (1 - s)
    }
    h = fma(S.a3p6, u, S.a2p4);
    G = fma(S.cp, rs, fma(u, h, S.a1p2));  // c'/s + d(poly u)/du * 2
    xd = fma(X, dx, Z * dz);               // (du/dt) / 2
    fp = fma(-G, xd, -dy);
}
template <bool GENERIC>
__device__ __forceinline__ void surf_eval(const SurfC &S, const bool refine, double Px, double Pz,
                                          double dx, double dy, double dz, double y0PyK, double t,
                                          double &f, double &fp, double &X, double &Z, double &G,
                                          double &w) {
    double u, rs, xd, h;
    surf_eval<GENERIC>(S, refine, Px, Pz, dx, dy, dz, y0PyK, t, f, fp, X, Z, G, w, u, rs, xd, h);
}

// fp32 pre-iterations: same two forms (REFINE surfaces keep the division, where the (1 - s)
// cancellation would cost |cK| * 2^-22 mm).  y0PyK = y0Py - cK for the division-free form.
template <bool DIVFREE>
__device__ __forceinline__ void surf_eval_f32(const SurfF &S, float Px, float Pz, float dx,
                                              float dy, float dz, float y0PyK, float t,
                                              float &f, float &fp) {
    const float X = fmaf(dx, t, Px);
    const float Z = fmaf(dz, t, Pz);
    const float u = fmaf(X, X, Z * Z);
    const float w = fmaf(-S.Kp, u, 1.0f);
    const float rs = rsqrt_f32(w);
    const float poly = fmaf(fmaf(S.a3p, u, S.a2p), u, S.a1p);
    const float Gp = fmaf(u, fmaf(S.a3p6, u, S.a2p4), S.a1p2);
    const float base = fmaf(-u, poly, fmaf(-dy, t, y0PyK));
    if (DIVFREE)
        f = fmaf(S.cK * w, rs, base);
    else
        f = fmaf(-u, S.cp * rcp_f32(fmaf(w, rs, 1.0f)), base);
    const float G = fmaf(S.cp, rs, Gp);
    fp = fmaf(-G, fmaf(X, dx, Z * dz), -dy);
}
// The same with the second derivative along the ray, for Halley steps (cubic convergence, no
// further MUFU): f'' = -(G (dx^2 + dz^2) + dG/du (du/dt)^2 / 2), dd = dx^2 + dz^2.
template <bool DIVFREE>
__device__ __forceinline__ float halley_step_f32(const SurfF &S, float Px, float Pz, float dx,
                                                 float dy, float dz, float dd, float y0PyK,
                                                 float t) {
    const float X = fmaf(dx, t, Px);
    const float Z = fmaf(dz, t, Pz);
    const float u = fmaf(X, X, Z * Z);
    const float w = fmaf(-S.Kp, u, 1.0f);
    const float rs = rsqrt_f32(w);
    const float poly = fmaf(fmaf(S.a3p, u, S.a2p), u, S.a1p);
    const float h = fmaf(S.a3p6, u, S.a2p4);
    const float Gp = fmaf(u, h, S.a1p2);
    const float base = fmaf(-u, poly, fmaf(-dy, t, y0PyK));
    float f;
    if (DIVFREE)
        f = fmaf(S.cK * w, rs, base);
    else
        f = fmaf(-u, S.cp * rcp_f32(fmaf(w, rs, 1.0f)), base);
    const float G = fmaf(S.cp, rs, Gp);
    const float xd = fmaf(X, dx, Z * dz);
    const float fp = fmaf(-G, xd, -dy);
    const float dGdu = fmaf(S.cKh * (rs * rs), rs, fmaf(S.a3p6, u, h));
    const float nfpp = fmaf(G, dd, 2.0f * dGdu * (xd * xd));        // -f''
    // t - 2 f f' / (2 f'^2 - f f'')
    const float den = fmaf(fp + fp, fp, f * nfpp);
    return fmaf(-(f + f) * fp, rcp_f32(den), t);
}

// Vector Snell refraction, ray_trace.py:27-35 with N x (-N x s) = s - N (N.s) and
// |N x s|^2 = 1 - (N.s)^2:  s2 = mu s + (mu cos_i - sqrt(1 - mu^2 (1 - cos_i^2))) N.
// Written on the UNNORMALISED normal n = (-gx, -1, -gz), |n|^2 = L2, m = -n.s = |n| cos_i:
//     s2 = mu s + n (mu m - sqrt((1 - mu^2) L2 + mu^2 m^2)) / L2
// which needs one reciprocal and one square root instead of two reciprocal square roots and
// drops the three multiplications that normalise N.  A negative radicand (total internal
// reflection) gives NaN, as in the reference.  (gx, gz, L2) are returned so that the unit
// normal of the last surface (out rows 5-7) can be formed once at the end.
__device__ __forceinline__ void refract(const SurfC &S, double X, double Z, double G,
                                        double &dx, double &dy, double &dz, double &gx,
                                        double &gz, double &L2) {
    gx = G * X;
    gz = G * Z;
    L2 = fma(gx, gx, fma(gz, gz, 1.0));
    const double m = fma(gx, dx, fma(gz, dz, dy));
    const double D = fma(S.mu2, m * m, S.omm2 * L2);
#if SOSAT_REFRACT_JOINT
    // one fully refined rsqrt serves both sqrt(D)/L2 and 1/L2: y = 1/(sqrt(D) L2),
    // sqrt(D)/L2 = D y, 1/L2 = (D y)(L2 y)
    const double y = SOSAT_RSQRT_SNELL(D * (L2 * L2));
    const double Dy = D * y;
    const double beta = fma(S.mu * m, Dy * (L2 * y), -Dy);
#else
    const double sqrtD = D * SOSAT_RSQRT_SNELL(D);
    const double beta = fma(S.mu, m, -sqrtD) * rcp_full(L2);
#endif
    dx = fma(S.mu, dx, -(gx * beta));
    dy = fma(S.mu, dy, -beta);
    dz = fma(S.mu, dz, -(gz * beta));
}
// local normal (-gx, -gy, 1)/|.| maps to tele (-gx, -1, -gz)/|.| (ray_trace.py:1162-1166)
__device__ __forceinline__ void unit_normal(double gx, double gz, double L2, double &nx,
                                            double &ny, double &nz) {
    const double rn = rsqrt_full(L2);
    nx = -gx * rn;
    ny = -rn;
    nz = -gz * rn;
}

// Stop / source planes, OPL total and validity: ray_trace.py:1510-1526, 1565-1580.
__device__ __forceinline__ void finish_ray(const GeomC &g, double Px, double Py, double Pz,
                                           double dx, double dy, double dz, double opl,
                                           double r2a_sq, RayResult &r) {
    const double idy = rcp_full(dy);
    const double t1 = fabs((g.y_lyot - Py) * idy);
    const double Lx = fma(dx, t1, Px), Ly = fma(dy, t1, Py), Lz = fma(dz, t1, Pz);
    const double t2 = fabs((g.y_source - Ly) * idy);
    r.ax = fma(dx, t2, Lx);
    r.ay = fma(dy, t2, Ly);
    r.az = fma(dz, t2, Lz);
    r.opl = opl + t1 + t2;
    r.dx = dx;
    r.dy = dy;
    r.dz = dz;
    // NaN anywhere makes both comparisons false, as in the reference
    r.valid = (fma(Lx, Lx, Lz * Lz) <= g.clip_stop_sq) && (r2a_sq <= g.clip_stop_sq) &&
              (r.opl == r.opl);
}

// The six-surface loop is unrolled so that the per-surface constants become immediate c[][]
// operands of the DFMAs and the si == 0 / si == 3 / conic-edge tests fold away (+4 % over the
// rolled loop with its LDC fetches, measured); -DSOSAT_UNROLL_SURFACES=0 keeps it rolled.
#if !defined(SOSAT_UNROLL_SURFACES) || SOSAT_UNROLL_SURFACES
#define SOSAT_SURFACE_LOOP_PRAGMA _Pragma("unroll")
#else
#define SOSAT_SURFACE_LOOP_PRAGMA _Pragma("unroll 1")
#endif

// Newton path.  N32 Newton steps in fp32 from the vertex-plane guess, then N64 in fp64.
// REFINE_MASK: bit si set = surface si uses the fully refined rsqrt / the division form of the
// fp32 pre-iterations (|c'/K'| > 300 mm).  EDGE_MASK: bit si set = surface si may have a
// bounded conic domain (K' > 0) and gets the conic-edge test.  Both are compile-time so that
// the unrolled surface loop holds one straight-line variant per surface (the instruction
// cache is a measured limiter of K1): the host launches the (0x04, 0x10) build when exactly
// lens 2b / lens 1b need them (the SAT prescription) and the (0x3f, 0x3f) build -- a few per
// cent slower -- for any other geometry.
// ADAPTIVE = false (fast path): a fixed schedule and nothing else; a lane whose last step is
// still above REFINE raises RAY_UNCONVERGED and the caller re-traces it with ADAPTIVE = true
// (out of line: essentially never taken for the boresight bundle, so the refinement loops
// stay out of the hot instruction stream).  ADAPTIVE = true: up to SOSAT_NEWTON_EXTRA more
// fp64 steps per surface; a lane still above TOL (or NaN) is marked for the bracketed solve.
// State of one ray between surfaces.
struct RayState {
    double Px, Py, Pz, dx, dy, dz;  // last hit point (or the feed) and direction
    double opl, r2a_sq;             // optical path so far; x^2 + z^2 at lens 2a (stop clip)
    double gx, gz, L2;              // unnormalised normal of the last surface
    unsigned flags;
    int step_hi;                    // high word of the largest |fp64 step| (N64 == 1 schedule)
};
__device__ __forceinline__ void ray_start(RayState &R, double Px, double Py, double Pz, double dx,
                                          double dy, double dz) {
    R.Px = Px; R.Py = Py; R.Pz = Pz;
    R.dx = dx; R.dy = dy; R.dz = dz;
    R.opl = 0.0; R.r2a_sq = 0.0;
    R.gx = 0.0; R.gz = 0.0; R.L2 = 1.0;
    R.flags = 0u;
    R.step_hi = 0;
}

#define SOSAT_SCHED_TPL \
    template <int N32, int N64, bool GENERIC, int REFINE_MASK, int EDGE_MASK, bool ADAPTIVE>
#define SOSAT_SCHED_ARGS N32, N64, GENERIC, REFINE_MASK, EDGE_MASK, ADAPTIVE

// y0 - Py (for the fp32 division form and the vertex-plane guess) and y0 - cK - Py.  The
// division-free surfaces of the default schedule ("lean") carry only the latter: one DADD.
SOSAT_SCHED_TPL __device__ __forceinline__ void surface_offsets(const GeomC &g, const int si,
                                                                const RayState &R, double &y0Py,
                                                                double &y0PyK) {
    constexpr bool kLinear = N32 > 0 && N64 == 1 && !ADAPTIVE;
    const SurfC &S = g.s[si];
    const bool lean = kLinear && !GENERIC && !((REFINE_MASK >> si) & 1);
    y0Py = lean ? 0.0 : S.y0 - R.Py;
    y0PyK = lean ? S.y0K - R.Py : (GENERIC ? y0Py : y0Py - S.cK);
}

// Stage 1 of a surface: the root parameter t from the fp32 steps (or the vertex-plane guess).
SOSAT_SCHED_TPL __device__ __forceinline__ double surface_seed(const GeomC &g, const int si,
                                                               const RayState &R) {
    // N64 == 1: Newton + Halley in fp32, one fp64 Newton step, slope factor carried to first order
    constexpr bool kLinear = N32 > 0 && N64 == 1 && !ADAPTIVE;
    const bool lean = kLinear && !GENERIC && !((REFINE_MASK >> si) & 1);
    double y0Py, y0PyK;
    surface_offsets<SOSAT_SCHED_ARGS>(g, si, R, y0Py, y0PyK);
    if (N32 == 0) return y0Py * rcp_half(R.dy);  // vertex-plane guess
    const SurfF &F = g.f[si];
    const float fPx = d2f_seed(R.Px), fPz = d2f_seed(R.Pz), fdx = d2f_seed(R.dx),
                fdy = d2f_seed(R.dy), fdz = d2f_seed(R.dz);
    const float fyK = lean ? d2f_seed(y0PyK) : 0.f;
    const float fy0 = lean ? fyK + F.cK : d2f_seed(y0Py);
    float tf = fy0 * rcp_f32(fdy);  // vertex-plane guess
    if (kLinear) {
        const float fdd = fmaf(fdx, fdx, fdz * fdz);
        if (!GENERIC && !((REFINE_MASK >> si) & 1)) {
            const float fy0K = lean ? fyK : fy0 - F.cK;
#pragma unroll
            for (int it = 0; it < N32 - 1; ++it) {
                float f, fp;
                surf_eval_f32<true>(F, fPx, fPz, fdx, fdy, fdz, fy0K, tf, f, fp);
                tf = fmaf(-f, rcp_f32(fp), tf);
            }
            tf = halley_step_f32<true>(F, fPx, fPz, fdx, fdy, fdz, fdd, fy0K, tf);
        } else {
#pragma unroll
            for (int it = 0; it < N32 - 1; ++it) {
                float f, fp;
                surf_eval_f32<false>(F, fPx, fPz, fdx, fdy, fdz, fy0, tf, f, fp);
                tf = fmaf(-f, rcp_f32(fp), tf);
            }
            tf = halley_step_f32<false>(F, fPx, fPz, fdx, fdy, fdz, fdd, fy0, tf);
        }
    } else if (!GENERIC && !((REFINE_MASK >> si) & 1)) {
        const float fy0K = fy0 - F.cK;
#pragma unroll
        for (int it = 0; it < N32; ++it) {
            float f, fp;
            surf_eval_f32<true>(F, fPx, fPz, fdx, fdy, fdz, fy0K, tf, f, fp);
            tf = fmaf(-f, rcp_f32(fp), tf);
        }
    } else {
#pragma unroll
        for (int it = 0; it < N32; ++it) {
            float f, fp;
            surf_eval_f32<false>(F, fPx, fPz, fdx, fdy, fdz, fy0, tf, f, fp);
            tf = fmaf(-f, rcp_f32(fp), tf);
        }
    }
    return f2d_seed(tf);
}

// Stage 2 of a surface: fp64 step(s) from the seed, hit point, OPL, refraction.
// t: in = the seed, out = the root parameter (kept by the caller as the seed history of the lane).
SOSAT_SCHED_TPL __device__ __forceinline__ void surface_hit(const GeomC &g, const int si,
                                                            RayState &R, double &t) {
    constexpr bool kLinear = N32 > 0 && N64 == 1 && !ADAPTIVE;
    const SurfC &S = g.s[si];
    const bool refine = ((REFINE_MASK >> si) & 1) != 0;
    double y0Py, y0PyK;
    surface_offsets<SOSAT_SCHED_ARGS>(g, si, R, y0Py, y0PyK);
    double X, Z, G, w, step = 0.0;
    if (kLinear) {
        // The fp32 root is converged to the fp32 floor (~1e-4 mm), so one fp64 Newton step
        // lands on the fp64 root and the slope factor and the hit coordinates are carried
        // to the new t to first order instead of by a second evaluation (9 DP instead of
        // 24 + 2 MUFU): dG/dt = dG/du du/dt, dG/du = 4 a2' + 12 a3' u + c'K'/(2 s^3)
        double f, fp, u, rs, xd, h;
        surf_eval<GENERIC>(S, refine, R.Px, R.Pz, R.dx, R.dy, R.dz, y0PyK, t, f, fp, X, Z, G, w, u, rs, xd, h);
        step = f * SOSAT_LINEAR_RCP(fp);  // the seed's 3e-5 relative error stays in t: <= 3e-8 mm
        t -= step;
        // 4 a2' + 12 a3' u = h + 6 a3' u with h = 4 a2' + 6 a3' u from the evaluation of G
        const double dGdu = fma(S.cKh, rs * rs * rs, fma(S.a3p6, u, h));
        G = fma(dGdu * xd, -2.0 * step, G);
        // (X, Z at the new t are the advanced hit point below: R.Px, R.Pz)
        // largest |step| of the ray, tracked on the high word (2 ALU instructions per surface
        // instead of a DSETP on the fp64 pipe) and tested once after the last surface
        R.step_hi = max(R.step_hi, __double2hiint(step) & 0x7fffffff);
    } else {
#pragma unroll
        for (int it = 0; it < N64; ++it) {
            double f, fp;
            surf_eval<GENERIC>(S, refine, R.Px, R.Pz, R.dx, R.dy, R.dz, y0PyK, t, f, fp, X, Z, G, w);
            step = f * rcp_seed(fp);
            t -= step;
        }
    }
    if (ADAPTIVE) {
#pragma unroll 1
        for (int extra = 0; extra < SOSAT_NEWTON_EXTRA && fabs(step) > SOSAT_NEWTON_REFINE;
             ++extra) {
            double f, fp;
            surf_eval<GENERIC>(S, refine, R.Px, R.Pz, R.dx, R.dy, R.dz, y0PyK, t, f, fp, X, Z, G, w);
            step = f * rcp_seed(fp);
            t -= step;
        }
        if (!(fabs(step) <= SOSAT_NEWTON_TOL)) R.flags |= RAY_FLAGGED;
    } else if (!kLinear) {
        if (fabs(step) > SOSAT_NEWTON_REFINE) R.flags |= RAY_UNCONVERGED;  // false for NaN
    }
    if (((EDGE_MASK >> si) & 1) && (EDGE_MASK == 0x10 || S.Kp > 0.0) && !(w >= SOSAT_EDGE_W))
        R.flags |= RAY_FLAGGED;
    // hit point (ray_trace.py:1133-1136 and equivalents)
    R.Px = fma(R.dx, t, R.Px);
    R.Py = fma(R.dy, t, R.Py);
    R.Pz = fma(R.dz, t, R.Pz);
    // lens-3 clip (ray_trace.py:1138-1140): the reference skips the ray; here the lane keeps
    // tracing so that the warp stays converged, and the result is discarded at the end
    if (si == 0 && fma(R.Px, R.Px, R.Pz * R.Pz) >= g.clip_l3_sq) R.flags |= RAY_CLIPPED_L3;
    if (si == 3) R.r2a_sq = fma(R.Px, R.Px, R.Pz * R.Pz);
    R.opl = fma(t, S.n_in, R.opl);  // dist_* = |hit - previous| = t for unit d
    if (kLinear)
        refract(S, R.Px, R.Pz, G, R.dx, R.dy, R.dz, R.gx, R.gz, R.L2);
    else
        refract(S, X, Z, G, R.dx, R.dy, R.dz, R.gx, R.gz, R.L2);
}

SOSAT_SCHED_TPL __device__ __forceinline__ void ray_finish(const GeomC &g, RayState &R,
                                                           RayResult &r) {
    constexpr bool kLinear = N32 > 0 && N64 == 1 && !ADAPTIVE;
    if (kLinear) {
        // |step| > MAX on some surface; NaN (exponent all ones) is not "unconverged", as before
        constexpr int kMaxHi = (int)(0x3ff00000u - ((unsigned)SOSAT_LINEAR_MAX_STEP_LOG2 << 20));
        if (R.step_hi >= kMaxHi && R.step_hi < 0x7ff00000) R.flags |= RAY_UNCONVERGED;
    }
    unit_normal(R.gx, R.gz, R.L2, r.nx, r.ny, r.nz);
    finish_ray(g, R.Px, R.Py, R.Pz, R.dx, R.dy, R.dz, R.opl, R.r2a_sq, r);
    r.flags = R.flags;
    if (R.flags & RAY_CLIPPED_L3) {
        r.flags = RAY_CLIPPED_L3;
        r.valid = false;
    }
}

SOSAT_SCHED_TPL __device__ __forceinline__ void trace_ray_newton(const GeomC &g, double Px, double Py,
                                                                 double Pz, double dx, double dy,
                                                                 double dz, RayResult &r) {
    RayState R;
    ray_start(R, Px, Py, Pz, dx, dy, dz);
    SOSAT_SURFACE_LOOP_PRAGMA
    for (int si = 0; si < SOSAT_NUM_SURFACES; ++si) {
        double t = surface_seed<SOSAT_SCHED_ARGS>(g, si, R);
        surface_hit<SOSAT_SCHED_ARGS>(g, si, R, t);
    }
    ray_finish<SOSAT_SCHED_ARGS>(g, R, r);
}

// The same, recording the six root parameters in hist[si * stride] (the seed history of the
// lane, shared memory: see trace_ray_predicted).
SOSAT_SCHED_TPL __device__ __forceinline__ void trace_ray_newton_hist(
    const GeomC &g, double Px, double Py, double Pz, double dx, double dy, double dz, RayResult &r,
    double *hist, const int stride) {
    RayState R;
    ray_start(R, Px, Py, Pz, dx, dy, dz);
    SOSAT_SURFACE_LOOP_PRAGMA
    for (int si = 0; si < SOSAT_NUM_SURFACES; ++si) {
        double t = surface_seed<SOSAT_SCHED_ARGS>(g, si, R);
        surface_hit<SOSAT_SCHED_ARGS>(g, si, R, t);
        hist[si * stride] = t;
    }
    ray_finish<SOSAT_SCHED_ARGS>(g, R, r);
}

// Predicted seeds.  In the fused kernel a CTA walks a strip of ADJACENT tiles of the launch grid,
// so the ray a lane traces in tile x is the neighbour, 16 launch columns on, of the ray it traced
// in tile x - 1.  The root parameter t_k of every surface is a smooth function of the launch
// angle, so its value in this tile is predicted from the lane's own last two tiles,
//     t_k(x) ~ 2 t_k(x - 1) - t_k(x - 2)        error (16 dtheta)^2 |t_k''|,
// <= 4e-4 mm at n_scan = 31623 (measured on the SAT prescription, worst on the 530 mm segment to
// lens 1b) -- the accuracy the fp32 Newton + Halley pre-iterations reach (1e-5 .. 3e-4 mm) for
// one DFMA and two shared-memory loads instead of ~70 instructions, 7 F2F and 4 MUFU per surface.
// The single fp64 Newton step and its first-order slope update are those of the "N32 + 1"
// schedule; a lane whose step exceeds 2^-10 mm (or is NaN: the history may be stale across the
// edge of the conic domain) raises RAY_UNCONVERGED and is re-traced by the adaptive path.
// h1 / h2: the lane's histories t(x-1), t(x-2), stride apart per surface; the new roots
// overwrite h2 (the caller swaps the roles).
template <bool GENERIC, int REFINE_MASK, int EDGE_MASK>
__device__ __forceinline__ void trace_ray_predicted(const GeomC &g, double Px, double Py, double Pz,
                                                    double dx, double dy, double dz, RayResult &r,
                                                    const double *h1, double *h2, const int stride) {
    RayState R;
    ray_start(R, Px, Py, Pz, dx, dy, dz);
    SOSAT_SURFACE_LOOP_PRAGMA
    for (int si = 0; si < SOSAT_NUM_SURFACES; ++si) {
        double t = fma(2.0, h1[si * stride], -h2[si * stride]);
        surface_hit<2, 1, GENERIC, REFINE_MASK, EDGE_MASK, false>(g, si, R, t);
        h2[si * stride] = t;
    }
    // NaN steps count as unconverged here (in the fp32-seeded schedule a NaN is the ray's own)
    constexpr int kMaxHi = (int)(0x3ff00000u - ((unsigned)SOSAT_LINEAR_MAX_STEP_LOG2 << 20));
    if (R.step_hi >= kMaxHi) R.flags |= RAY_UNCONVERGED;
    ray_finish<2, 1, GENERIC, REFINE_MASK, EDGE_MASK, false>(g, R, r);
}

// ---- slow path: the reference's bracketed solve ---------------------------------------------
// root_z?? with its "sag NaN -> 0" rule (ray_trace.py:1125-1126): outside the conic domain
// the function degenerates to the vertex plane.
__device__ __forceinline__ double root_fn_ref(const SurfC &S, double Px, double Pz, double dx,
                                              double dy, double dz, double y0Py, double t) {
    const double X = fma(dx, t, Px), Z = fma(dz, t, Pz);
    const double u = fma(X, X, Z * Z);
    const double w = fma(-S.Kp, u, 1.0);
    double sag = 0.0;
    if (w >= 0.0) {
        const double s = sqrt(w);
        sag = u * (S.cp / (1.0 + s) + fma(fma(S.a3p, u, S.a2p), u, S.a1p));
    }
    return fma(-dy, t, y0Py) - sag;
}

// Brent's method as published (Brent 1973) and as scipy.optimize.brentq applies it with its
// defaults xtol = 2e-12, rtol = 4 eps, maxiter = 100; the reference calls it with exactly
// those defaults.  Returns NaN when f is NaN at a bracket end or the signs agree.
__device__ __noinline__ double brent_solve(const SurfC &S, double Px, double Pz, double dx,
                                           double dy, double dz, double y0Py,
                                           unsigned &flags) {
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    double xpre = S.t_lo, xcur = S.t_hi;
    double xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
    double fpre = root_fn_ref(S, Px, Pz, dx, dy, dz, y0Py, xpre);
    double fcur = root_fn_ref(S, Px, Pz, dx, dy, dz, y0Py, xcur);
    if (fpre != fpre || fcur != fcur) return nan("");
    if (fpre == 0.0) return xpre;
    if (fcur == 0.0) return xcur;
    if ((fpre < 0.0) == (fcur < 0.0)) {
        flags |= RAY_BRACKET_FAIL;
        return nan("");
    }
    for (int i = 0; i < 100; ++i) {
        if (fpre != 0.0 && fcur != 0.0 && ((fpre < 0.0) != (fcur < 0.0))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0.0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur;
                scur = stry;
            } else {
                spre = sbis;
                scur = sbis;
            }
        } else {
            spre = sbis;
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = root_fn_ref(S, Px, Pz, dx, dy, dz, y0Py, xcur);
        if (fcur != fcur) return nan("");
    }
    return xcur;
}

__device__ __noinline__ void trace_ray_bracketed(const GeomC &g, double Px, double Py,
                                                 double Pz, double dx, double dy, double dz,
                                                 RayResult &r) {
    double opl = 0.0, r2a_sq = 0.0;
    double gx = 0.0, gz = 0.0, L2 = 1.0;
    unsigned flags = 0;
#pragma unroll 1
    for (int si = 0; si < SOSAT_NUM_SURFACES; ++si) {
        const SurfC &S = g.s[si];
        const double y0Py = S.y0 - Py;
        const double t = brent_solve(S, Px, Pz, dx, dy, dz, y0Py, flags);
        double f, fp, X, Z, G, w;
        surf_eval<true>(S, false, Px, Pz, dx, dy, dz, y0Py, t, f, fp, X, Z, G, w);  // slope at the hit
        Px = fma(dx, t, Px);
        Py = fma(dy, t, Py);
        Pz = fma(dz, t, Pz);
        if (si == 0 && fma(Px, Px, Pz * Pz) >= g.clip_l3_sq) {
            r.flags = RAY_CLIPPED_L3;
            r.valid = false;
            return;
        }
        if (si == 3) r2a_sq = fma(Px, Px, Pz * Pz);
        opl = fma(t, S.n_in, opl);
        refract(S, X, Z, G, dx, dy, dz, gx, gz, L2);
    }
    unit_normal(gx, gz, L2, r.nx, r.ny, r.nz);
    r.flags = flags;
    finish_ray(g, Px, Py, Pz, dx, dy, dz, opl, r2a_sq, r);
}

// Out-of-line adaptive Newton path (pure fp64 start, 4 fixed steps + refinement).
template <bool GENERIC, int REFINE_MASK, int EDGE_MASK>
__device__ __noinline__ void trace_ray_adaptive(const GeomC &g, double Px, double Py, double Pz,
                                                double dx, double dy, double dz, RayResult &r) {
    trace_ray_newton<0, 4, GENERIC, REFINE_MASK, EDGE_MASK, true>(g, Px, Py, Pz, dx, dy, dz, r);
}

// Fast path; the adaptive path for lanes the fixed schedule left unconverged; the bracketed
// slow path for rays that came out valid but were flagged.
template <int N32, int N64, bool GENERIC, int REFINE_MASK, int EDGE_MASK>
__device__ __forceinline__ void trace_ray(const GeomC &g, double Px, double Py, double Pz,
                                          double dx, double dy, double dz, RayResult &r,
                                          unsigned &n_slow) {
    trace_ray_newton<N32, N64, GENERIC, REFINE_MASK, EDGE_MASK, false>(g, Px, Py, Pz, dx, dy, dz, r);
    if (r.flags & RAY_UNCONVERGED)
        trace_ray_adaptive<GENERIC, REFINE_MASK, EDGE_MASK>(g, Px, Py, Pz, dx, dy, dz, r);
    if (r.valid && (r.flags & RAY_FLAGGED)) {
        trace_ray_bracketed(g, Px, Py, Pz, dx, dy, dz, r);
        r.flags |= RAY_FLAGGED;
        ++n_slow;
    }
}

}  // namespace sosat
