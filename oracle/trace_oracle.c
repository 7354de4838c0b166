/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the reference ray tracer.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (sosat_optics_b200) never does.
 *
 * Restates, in plain C double arithmetic and in the reference's own operation order:
 *   sosat_optics/ot_geo.py:13-14,32-46      constants (passed in by the caller)
 *   sosat_optics/ot_geo.py:49-91,182-227,313-354     asphere sag  z{1,2,3}{a,b}
 *   sosat_optics/ot_geo.py:94-143,230-276,364-414    asphere slope d_z{1,2,3}{a,b}
 *   sosat_optics/ot_geo.py:162-179,295-310,437-458   tele_into_m* frame maps
 *   sosat_optics/ray_trace.py:27-35          snell_vec
 *   sosat_optics/ray_trace.py:1086-1103      launch grid
 *   sosat_optics/ray_trace.py:1114-1506      six refracting surfaces
 *   sosat_optics/ray_trace.py:1510-1532      Lyot / source planes, OPL, launch angles
 *   sosat_optics/ray_trace.py:1565-1588      validity mask + out[17, n] layout
 *
 * Third-party arithmetic on the path: scipy.optimize.brentq (not vendored by the
 * reference and not version-pinned; setup.py:32-40 declares no dependencies).  Its
 * published algorithm (Brent 1973 as implemented in scipy/optimize/Zeros/brentq.c,
 * defaults xtol=2e-12, rtol=4*DBL_EPSILON, maxiter=100 from scipy/optimize/_zeros_py.py)
 * is restated in brentq_restated() below and is checked against the installed
 * scipy.optimize.brentq in tests/test_oracle.py.
 *
 * Parity pin: the reference has no tests; this oracle is pinned against golden vectors
 * produced by running the reference itself (oracle/make_golden.py -> tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double c, k, a1, a2, a3; /* asphere coefficients, r in cm (ot_geo.py:57-61 etc.) */
    double y0;               /* vertex position on the tele y axis, mm */
    double th;               /* frame tilt angle (pi/2 for every surface, ot_geo.py:32-37) */
    double n1, n2;           /* refractive index before / after the surface */
    double t_lo, t_hi;       /* brentq bracket, mm (ray_trace.py:1130,1198,...) */
} oracle_surface_t;

typedef struct {
    oracle_surface_t surf[6]; /* trace order 3b,3a,2b,2a,1b,1a */
    double y_lyot;            /* ot_geo.py:46 */
    double y_source;          /* SatGeo.y_source */
    double fwhp_x, fwhp_y;    /* SatGeo.th_fwhp_x / th_fwhp_y, rad */
    double clip_l3;           /* 392/2 mm, ray_trace.py:1138 */
    double clip_stop;         /* 210 mm, ray_trace.py:1569-1570 */
    double cone;              /* 0.4 rad, ray_trace.py:1086-1087 */
} oracle_geom_t;

/* ---- ot_geo.py sag: amp*10 with r in cm ------------------------------------------ */
static double sag(const oracle_surface_t *s, double x, double y)
{
    double xt = x / 1e1, yt = y / 1e1;
    double r = sqrt(xt * xt + yt * yt);
    double r2 = r * r;
    double amp = (s->c * r2) / (1 + sqrt(1 - ((1 + s->k) * (s->c * s->c) * r2)));
    amp += s->a1 * r2 + s->a2 * pow(r, 4) + s->a3 * pow(r, 6);
    return amp * 10;
}

/* ---- ot_geo.py d_z*: analytic slope (dimensionless) -------------------------------- */
static void slope(const oracle_surface_t *s, double x, double y, double *gx, double *gy)
{
    double xt = x / 1e1, yt = y / 1e1;
    double r = sqrt(xt * xt + yt * yt);
    double r2 = r * r;
    double c = s->c, k = s->k;
    double root = sqrt(1 - ((1 + k) * (c * c) * r2));
    double coeff_1 = (s->a1 * 2) + (s->a2 * 4 * r2) + (s->a3 * 6 * pow(r, 4));
    double coeff_2 = (c * 2) / (1 + root);
    double coeff_3 = ((c * c * c) * (k + 1) * r2) / (root * ((1 + root) * (1 + root)));
    *gx = xt * (coeff_1 + coeff_2 + coeff_3);
    *gy = yt * (coeff_1 + coeff_2 + coeff_3);
}

/* ---- ot_geo.py tele_into_m*: translate along y, rotate by -th about x --------------- */
static void tele_into(const oracle_surface_t *s, double x, double y, double z,
                      double *xo, double *yo, double *zo)
{
    double cm = cos(-s->th), sm = sin(-s->th);
    y -= s->y0;
    *xo = x;
    *yo = y * cm - z * sm;
    *zo = y * sm + z * cm;
}

typedef struct {
    const oracle_surface_t *s;
    double o[3], d[3];
} root_ctx_t;

/* the root_z?? closures, e.g. ray_trace.py:1114-1128 (sag NaN -> 0) */
static double root_fn(const root_ctx_t *q, double t)
{
    double x = q->o[0] + q->d[0] * t;
    double y = q->o[1] + q->d[1] * t;
    double z = q->o[2] + q->d[2] * t;
    double xm, ym, zm;
    tele_into(q->s, x, y, z, &xm, &ym, &zm);
    double zs = sag(q->s, xm, ym);
    if (isnan(zs)) zs = 0;
    return zm - zs;
}

/* status: 0 ok, 1 sign error (reference would raise ValueError), 2 NaN function value
 * (old SciPy: garbage root, ray later zeroed; oracle/ref_shim.py returns NaN), 3 no conv. */
static double brentq_restated(const root_ctx_t *q, double xa, double xb, int *status,
                              long *funcalls)
{
    const double xtol = 2e-12, rtol = 4 * 2.220446049250313e-16;
    const int maxiter = 100;
    double xpre = xa, xcur = xb;
    double xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    *status = 0;
    fpre = root_fn(q, xpre);
    fcur = root_fn(q, xcur);
    *funcalls += 2;
    if (isnan(fpre) || isnan(fcur)) { *status = 2; return NAN; }
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) { *status = 1; return NAN; }
    for (int i = 0; i < maxiter; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre); /* secant */
            } else {
                dpre = (fpre - fcur) / (xpre - xcur); /* inverse quadratic */
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double lim = fabs(spre);
            double lim2 = 3 * fabs(sbis) - delta;
            if (lim2 < lim) lim = lim2;
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = root_fn(q, xcur);
        *funcalls += 1;
        if (isnan(fcur)) { *status = 2; return NAN; }
    }
    *status = 3;
    return xcur;
}

static void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* ray_trace.py:27-35 -- sqrt of a negative (total internal reflection) gives NaN */
static void snell_vec(double n1, double n2, const double *N, const double *s1, double *s2)
{
    double negN[3] = {-N[0], -N[1], -N[2]};
    double c1[3], c2[3], c3[3];
    cross3(negN, s1, c1);
    cross3(N, c1, c2);
    cross3(N, s1, c3);
    double mu = n1 / n2;
    double rad = sqrt(1 - (mu * mu) * (c3[0] * c3[0] + c3[1] * c3[1] + c3[2] * c3[2]));
    for (int i = 0; i < 3; i++) s2[i] = mu * c2[i] - N[i] * rad;
}

/* numpy.linspace(start, stop, n)[i]: arange(n)*step + start, last element forced to stop */
static double linspace_at(double start, double stop, long n, long i)
{
    if (n == 1) return start;
    double step = (stop - start) / (double)(n - 1);
    if (i == n - 1) return stop;
    return (double)i * step + start;
}

/* One ray, index ii = iphi*n_scan + itheta (ray_trace.py:1089-1091).  col = 17 doubles.
 * Returns brentq status of the first failing solve (0 if all fine). */
static int trace_one(const oracle_geom_t *g, const double *rx, long n_scan, long ii,
                     double *col, long *funcalls)
{
    long ith = ii % n_scan, iph = ii / n_scan;
    const double half_pi = 1.5707963267948966;
    double th = linspace_at(half_pi - g->cone, half_pi + g->cone, n_scan, ith);
    double ph = linspace_at(half_pi - g->cone, half_pi + g->cone, n_scan, iph);
    double P[3] = {rx[0], rx[1], rx[2]};
    double d[3] = {sin(th) * cos(ph), sin(th) * sin(ph), cos(th)};
    double dist[6], tan_first_t[3] = {0, 0, 0}, N_hat_t[3] = {0, 0, 0};
    double p2a[3] = {0, 0, 0};
    int first_bad = 0;
    memset(col, 0, 17 * sizeof(double));

    for (int si = 0; si < 6; si++) {
        const oracle_surface_t *s = &g->surf[si];
        root_ctx_t q;
        q.s = s;
        memcpy(q.o, P, sizeof P);
        memcpy(q.d, d, sizeof d);
        int st;
        double t = brentq_restated(&q, s->t_lo, s->t_hi, &st, funcalls);
        if (st && !first_bad) first_bad = st;
        double H[3] = {P[0] + d[0] * t, P[1] + d[1] * t, P[2] + d[2] * t};
        if (si == 0 && H[0] * H[0] + H[2] * H[2] >= g->clip_l3 * g->clip_l3)
            return first_bad; /* ray_trace.py:1138-1140: whole column stays zero */
        double hl[3], pl[3], gx, gy;
        tele_into(s, H[0], H[1], H[2], &hl[0], &hl[1], &hl[2]);
        tele_into(s, P[0], P[1], P[2], &pl[0], &pl[1], &pl[2]);
        slope(s, hl[0], hl[1], &gx, &gy);
        double nt[3] = {-gx, -gy, 1};
        double nn = sqrt(nt[0] * nt[0] + nt[1] * nt[1] + nt[2] * nt[2]);
        double N[3] = {nt[0] / nn, nt[1] / nn, nt[2] / nn};
        double v[3] = {hl[0] - pl[0], hl[1] - pl[1], hl[2] - pl[2]};
        dist[si] = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        double tin[3] = {v[0] / dist[si], v[1] / dist[si], v[2] / dist[si]};
        double tout[3];
        snell_vec(s->n1, s->n2, N, tin, tout);
        /* rotate back to the tele frame (e.g. ray_trace.py:1162-1179) */
        double ct = cos(s->th), sn = sin(s->th);
        N_hat_t[0] = N[0];
        N_hat_t[1] = N[1] * ct - N[2] * sn;
        N_hat_t[2] = N[1] * sn + N[2] * ct;
        if (si == 0) {
            tan_first_t[0] = tin[0];
            tan_first_t[1] = tin[1] * ct - tin[2] * sn;
            tan_first_t[2] = tin[1] * sn + tin[2] * ct;
        }
        d[0] = tout[0];
        d[1] = tout[1] * ct - tout[2] * sn;
        d[2] = tout[1] * sn + tout[2] * ct;
        memcpy(P, H, sizeof P);
        if (si == 3) memcpy(p2a, H, sizeof H);
    }
    /* ray_trace.py:1510-1526 */
    double dist_lyot = fabs((g->y_lyot - P[1]) / d[1]);
    double pos_lyot[3] = {P[0] + dist_lyot * d[0], P[1] + dist_lyot * d[1],
                          P[2] + dist_lyot * d[2]};
    double dist_ap = fabs((g->y_source - pos_lyot[1]) / d[1]);
    double n_si = g->surf[0].n2;
    double total = dist[0] + dist[1] * n_si + dist[2] + dist[3] * n_si + dist[4] +
                   dist[5] * n_si + dist_lyot + dist_ap;
    double pos_ap[3] = {pos_lyot[0] + dist_ap * d[0], pos_lyot[1] + dist_ap * d[1],
                        pos_lyot[2] + dist_ap * d[2]};
    /* ray_trace.py:1529-1532 */
    double de_ve = atan(tan_first_t[0] / (-tan_first_t[1]));
    double de_ho = atan(tan_first_t[2] /
                        sqrt(tan_first_t[0] * tan_first_t[0] + tan_first_t[1] * tan_first_t[1]));
    col[0] = pos_ap[0];
    col[1] = pos_ap[1];
    col[2] = pos_ap[2];
    double r2 = g->clip_stop * g->clip_stop;
    if ((pos_lyot[0] * pos_lyot[0] + pos_lyot[2] * pos_lyot[2] <= r2) &&
        (p2a[0] * p2a[0] + p2a[2] * p2a[2] <= r2)) {
        col[3] = total;
        double sig = 1 / sqrt(8 * log(2.0));
        double hx = de_ho / g->fwhp_x, hy = de_ve / g->fwhp_y;
        col[4] = exp((-0.5) * (hx * hx + hy * hy) / (sig * sig));
    } else {
        col[3] = 0;
        col[4] = 0;
    }
    col[5] = N_hat_t[0]; col[6] = N_hat_t[1]; col[7] = N_hat_t[2];
    col[8] = d[0]; col[9] = d[1]; col[10] = d[2];
    return first_bad;
}

/* out: [17][ray_end-ray_begin] row-major (reference layout ray_trace.py:1095).
 * status (may be NULL): per-ray brentq status.  stats[0] = total function evaluations.
 * n_threads <= 0 -> OpenMP default. Returns number of rays whose solve reported a sign error. */
long oracle_rx_to_lyot(const oracle_geom_t *g, const double *rx, long n_scan, long ray_begin,
                       long ray_end, double *out, int *status, long *stats, int n_threads)
{
    long n = ray_end - ray_begin;
    long sign_err = 0, calls = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : sign_err, calls)
    for (long j = 0; j < n; j++) {
        double col[17];
        long fc = 0;
        int st = trace_one(g, rx, n_scan, ray_begin + j, col, &fc);
        for (int r = 0; r < 17; r++) out[(long)r * n + j] = col[r];
        if (status) status[j] = st;
        if (st == 1) sign_err++;
        calls += fc;
    }
    if (stats) stats[0] = calls;
    return sign_err;
}

/* Standalone root find on one surface for testing the brentq restatement against SciPy. */
double oracle_brentq_surface(const oracle_surface_t *s, const double *o, const double *d,
                             double lo, double hi, int *status)
{
    root_ctx_t q;
    long fc = 0;
    q.s = s;
    memcpy(q.o, o, 3 * sizeof(double));
    memcpy(q.d, d, 3 * sizeof(double));
    return brentq_restated(&q, lo, hi, status, &fc);
}

double oracle_root_fn(const oracle_surface_t *s, const double *o, const double *d, double t)
{
    root_ctx_t q;
    q.s = s;
    memcpy(q.o, o, 3 * sizeof(double));
    memcpy(q.d, d, 3 * sizeof(double));
    return root_fn(&q, t);
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
