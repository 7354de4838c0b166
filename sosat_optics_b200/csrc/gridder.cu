// Scattered rays -> regular grid -> stage grid: the glue between stage 1 and stage 2 of the
// reference's notebooks, on the device.
//
//   sim2d     ray_trace.py:1675-1717  scipy.interpolate.griddata(method="cubic") of amplitude and
//                                     phase on a 100 x 100 grid over the valid rays within 21 cm
//                                     of the beam centre, zero beyond 20 cm
//   sim_data  ray_trace.py:1720-1754  interp2d(kind="cubic") resampling onto the stage grid
//
// Both call third-party code that the reference does not vendor (SciPy, unpinned): griddata's
// cubic method is CloughTocher2DInterpolator on Qhull's Delaunay triangulation with gradients
// from a global curvature-minimising Gauss-Seidel iteration (scipy/interpolate/interpnd.pyx:
// _estimate_gradients_2d_global, _clough_tocher_2d_single; Nielson 1983, Renka & Cline 1984);
// interp2d(kind="cubic") is FITPACK's interpolating bicubic spline with not-a-knot knots.  Their
// published algorithms are restated here for the one input this path ever sees: the ray
// positions are a smoothly WARPED STRUCTURED LATTICE (ray ii = iphi * n_scan + itheta), so
//   * the Delaunay triangulation is local: every lattice quad with four selected corners is cut
//     along the diagonal that passes the in-circle test, a quad with three selected corners is
//     one triangle (checked against scipy.spatial.Delaunay on the notebook bundle: 13 046 of
//     13 346 triangles identical; the rest are co-circular quads on the symmetry axes, where
//     either diagonal is Delaunay and Qhull's choice is roundoff, and the hull fill-in slivers
//     outside the lattice, all beyond the 20 cm zeroing radius);
//   * vertex neighbours, triangle neighbours and point location follow from lattice indices --
//     no Qhull, no search structure;
//   * the Gauss-Seidel sweeps run in a four-colour order (node (r, c) has colour (r&1, c&1); no
//     two neighbours share one), i.e. exact Gauss-Seidel in another vertex order, converged
//     further than scipy's 1e-6 stopping rule.
// Agreement with the reference's own output: tests/golden/farfield_cfg2_240.npz (n_scan = 100)
// amplitude grid 2.5e-5, stage-grid amplitude 1e-6, far field 6.5e-5 of peak, FWHM 21.5039 arcmin
// -- the conditioning of the reference's glue under a 1e-9 cm perturbation of the rays (DESIGN.md
// 9); tests/golden/farfield_cfg2loop_1000_90ghz.npz (n_scan = 150, 1000^2 far field) 8e-7 of peak.
#include <cmath>
#include <cstdint>

#include "common.cuh"

namespace sosat {

// scalars shared by the kernels of one sim2d call
enum { SC_CX = 0, SC_CY, SC_XMIN, SC_XMAX, SC_YMIN, SC_YMAX, SC_NSEL, SC_CENTRE, SC_SUMX, SC_SUMY, SC_BEST, SC_COUNT };

__device__ __forceinline__ void atomic_min_double(double *dst, double v) {
    unsigned long long *a = (unsigned long long *)dst, old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (old != assumed);
}
__device__ __forceinline__ void atomic_max_double(double *dst, double v) {
    unsigned long long *a = (unsigned long long *)dst, old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (old != assumed);
}

__global__ void g2d_init_kernel(double *sc, int *node, long long n_nodes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        for (int q = 0; q < SC_COUNT; ++q) sc[q] = 0.0;
        sc[SC_XMIN] = sc[SC_YMIN] = INFINITY;
        sc[SC_XMAX] = sc[SC_YMAX] = -INFINITY;
        sc[SC_BEST] = INFINITY;
    }
    for (long long k = i; k < n_nodes; k += (long long)gridDim.x * blockDim.x) node[k] = -1;
}

// get_beam_cent (ray_trace.py:1652-1654): mean of x_sim, y_sim over ALL points
__global__ void g2d_sum_kernel(const double *__restrict__ x, const double *__restrict__ y, long long n,
                               double *sc) {
    double sx = 0.0, sy = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        sx += x[i];
        sy += y[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sc + SC_SUMX, sx);
        atomicAdd(sc + SC_SUMY, sy);
    }
}

// selection (ray_trace.py:1681-1685), its bounding box (the extent of np.mgrid, :1686-1689), the
// node table of the lattice, and the selected point nearest to the centre (start of the walks)
__global__ void g2d_select_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                  const long long *__restrict__ ray_index, long long n, double r_sel2,
                                  double *sc, int *node) {
    const double cx = sc[SC_SUMX] / (double)n, cy = sc[SC_SUMY] / (double)n;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc[SC_CX] = cx;
        sc[SC_CY] = cy;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double px = x[i], py = y[i];
        const double r2 = (px - cx) * (px - cx) + (py - cy) * (py - cy);
        if (px == px && py == py && r2 <= r_sel2) {
            node[ray_index[i]] = (int)i;
            atomic_min_double(sc + SC_XMIN, px);
            atomic_max_double(sc + SC_XMAX, px);
            atomic_min_double(sc + SC_YMIN, py);
            atomic_max_double(sc + SC_YMAX, py);
            atomicAdd(sc + SC_NSEL, 1.0);
            atomic_min_double(sc + SC_BEST, r2);
        }
    }
}
__global__ void g2d_centre_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                  const long long *__restrict__ ray_index, long long n, double r_sel2,
                                  double *sc) {
    const double cx = sc[SC_CX], cy = sc[SC_CY], best = sc[SC_BEST];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double px = x[i], py = y[i];
        const double r2 = (px - cx) * (px - cx) + (py - cy) * (py - cy);
        if (r2 == best) sc[SC_CENTRE] = (double)ray_index[i];  // ties: any of them will do
    }
}

// ---- triangulation of the lattice ------------------------------------------------------------
// Quad (r, c) has the corners q0 = (r, c), q1 = (r, c+1), q2 = (r+1, c+1), q3 = (r+1, c).  Flag:
//   0      no triangle             1   diagonal q0-q2: (q0,q1,q2), (q0,q2,q3)
//   2      diagonal q1-q3: (q0,q1,q3), (q1,q2,q3)        3 + m   corner m missing: one triangle
struct Lattice {
    const double *x, *y;
    const int *node;
    const unsigned char *quad;
    int n;  // n_scan
    __device__ __forceinline__ int corner(int r, int c, int k) const {
        const int rr = r + (k >> 1), cc = c + ((k == 1 || k == 2) ? 1 : 0);
        return node[(long long)rr * n + cc];
    }
    __device__ __forceinline__ int flag(int r, int c) const {
        if (r < 0 || c < 0 || r >= n - 1 || c >= n - 1) return 0;
        return quad[(long long)r * (n - 1) + c];
    }
};
// corners (0..3) of triangle t of a quad with flag f; returns false if there is no such triangle
__device__ __forceinline__ bool tri_corners(int f, int t, int k[3]) {
    if (f == 1) {
        k[0] = 0; k[1] = t ? 2 : 1; k[2] = t ? 3 : 2;
        return t < 2;
    }
    if (f == 2) {
        k[0] = t ? 1 : 0; k[1] = t ? 2 : 1; k[2] = 3;
        return t < 2;
    }
    if (f >= 3 && t == 0) {
        const int m = f - 3;
        for (int q = 0, o = 0; q < 4; ++q)
            if (q != m) k[o++] = q;
        return true;
    }
    return false;
}

__global__ void g2d_quad_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                const int *__restrict__ node, int n, unsigned char *quad) {
    const long long nq = (long long)(n - 1) * (n - 1);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq;
         q += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(q / (n - 1)), c = (int)(q % (n - 1));
        const int v[4] = {node[(long long)r * n + c], node[(long long)r * n + c + 1],
                          node[(long long)(r + 1) * n + c + 1], node[(long long)(r + 1) * n + c]};
        const int present = (v[0] >= 0) + (v[1] >= 0) + (v[2] >= 0) + (v[3] >= 0);
        unsigned char f = 0;
        if (present == 4) {
            // in-circle test of q3 against the circle through q0, q1, q2 (orientation-normalised):
            // inside -> the diagonal q0-q2 is not Delaunay, use q1-q3
            const double dx = x[v[3]], dy = y[v[3]];
            const double ax = x[v[0]] - dx, ay = y[v[0]] - dy, bx = x[v[1]] - dx, by = y[v[1]] - dy,
                         cx = x[v[2]] - dx, cy = y[v[2]] - dy;
            const double ic = (ax * ax + ay * ay) * (bx * cy - cx * by) -
                              (bx * bx + by * by) * (ax * cy - cx * ay) +
                              (cx * cx + cy * cy) * (ax * by - bx * ay);
            const double orient = (x[v[1]] - x[v[0]]) * (y[v[2]] - y[v[0]]) -
                                  (y[v[1]] - y[v[0]]) * (x[v[2]] - x[v[0]]);
            f = (orient > 0 ? ic : -ic) > 0 ? 2 : 1;
        } else if (present == 3) {
            for (int m = 0; m < 4; ++m)
                if (v[m] < 0) f = (unsigned char)(3 + m);
        }
        quad[q] = f;
    }
}

// ---- gradient estimation: one Gauss-Seidel half-sweep over the nodes of one colour ------------
// scipy interpnd.pyx _estimate_gradients_2d_global: at every vertex minimise the sum over its
// edges of int |W''|^2 of the Clough-Tocher edge cubic with the neighbours' gradients fixed:
//   Q = 4 sum E E^T / L^3,  s = sum (6 (f1 - f2) - 2 df2) E / L^3,  df2 = -E . grad(V2),
//   grad(V1) = -Q^-1 s.   Two fields (amplitude, phase) share the geometry.
__global__ void g2d_gradient_kernel(Lattice L, const double *__restrict__ fa, const double *__restrict__ fp,
                                    double *grad /* [n_pts][4]: ax ay px py */, int colour) {
    const int n = L.n, half_r = (n - (colour >> 1) + 1) / 2, half_c = (n - (colour & 1) + 1) / 2;
    const long long total = (long long)half_r * half_c;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        const int r = 2 * (int)(k / half_c) + (colour >> 1), c = 2 * (int)(k % half_c) + (colour & 1);
        const int ip = L.node[(long long)r * n + c];
        if (ip < 0) continue;
        const int f00 = L.flag(r - 1, c - 1), f01 = L.flag(r - 1, c), f10 = L.flag(r, c - 1), f11 = L.flag(r, c);
        double Q0 = 0, Q1 = 0, Q3 = 0, sa0 = 0, sa1 = 0, sp0 = 0, sp1 = 0;
        const double px = L.x[ip], py = L.y[ip], a1 = fa[ip], p1 = fp[ip];
        int n_edges = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int dr = e < 4 ? (e == 2 ? 1 : e == 3 ? -1 : 0) : ((e & 2) ? 1 : -1);
            const int dc = e < 4 ? (e == 0 ? 1 : e == 1 ? -1 : 0) : ((e & 1) ? 1 : -1);
            const int r2 = r + dr, c2 = c + dc;
            if (r2 < 0 || c2 < 0 || r2 >= n || c2 >= n) continue;
            const int ip2 = L.node[(long long)r2 * n + c2];
            if (ip2 < 0) continue;
            bool edge;
            if (e < 4) {  // a lattice side: an edge iff one of the two quads sharing it is triangulated
                const int fa_ = dr == 0 ? (dc > 0 ? f01 : f00) : (dr > 0 ? f10 : f00);
                const int fb_ = dr == 0 ? (dc > 0 ? f11 : f10) : (dr > 0 ? f11 : f01);
                edge = fa_ != 0 || fb_ != 0;
            } else {      // a diagonal of the quad between the two nodes
                const int f = dr > 0 ? (dc > 0 ? f11 : f10) : (dc > 0 ? f01 : f00);
                const bool main_diag = (dr > 0) == (dc > 0);  // q0-q2 of that quad
                edge = main_diag ? (f == 1 || f == 4 || f == 6) : (f == 2 || f == 3 || f == 5);
            }
            if (!edge) continue;
            ++n_edges;
            const double ex = L.x[ip2] - px, ey = L.y[ip2] - py;
            const double len = sqrt(ex * ex + ey * ey), L3 = len * len * len;
            const double *g2 = grad + 4ll * ip2;
            const double dfa = -ex * g2[0] - ey * g2[1], dfp = -ex * g2[2] - ey * g2[3];
            Q0 += 4 * ex * ex / L3;
            Q1 += 4 * ex * ey / L3;
            Q3 += 4 * ey * ey / L3;
            const double wa = 6 * (a1 - fa[ip2]) - 2 * dfa, wp = 6 * (p1 - fp[ip2]) - 2 * dfp;
            sa0 += wa * ex / L3;
            sa1 += wa * ey / L3;
            sp0 += wp * ex / L3;
            sp1 += wp * ey / L3;
        }
        double *g1 = grad + 4ll * ip;
        const double det = Q0 * Q3 - Q1 * Q1;
        if (n_edges < 2 || det == 0.0) {  // isolated node / collinear edges: scipy divides by zero here
            g1[0] = g1[1] = g1[2] = g1[3] = 0.0;
            continue;
        }
        g1[0] = -(Q3 * sa0 - Q1 * sa1) / det;
        g1[1] = -(-Q1 * sa0 + Q0 * sa1) / det;
        g1[2] = -(Q3 * sp0 - Q1 * sp1) / det;
        g1[3] = -(-Q1 * sp0 + Q0 * sp1) / det;
    }
}

// ---- evaluation -------------------------------------------------------------------------------
struct Tri {
    int v[3];       // point indices
    int r, c, t;    // quad and triangle number inside it
    int k[3];       // corners of the quad
};
__device__ __forceinline__ bool load_tri(const Lattice &L, int r, int c, int t, Tri &T) {
    const int f = L.flag(r, c);
    if (!tri_corners(f, t, T.k)) return false;
    T.r = r; T.c = c; T.t = t;
    for (int q = 0; q < 3; ++q) T.v[q] = L.corner(r, c, T.k[q]);
    return true;
}
// barycentric coordinates of (X, Y) in triangle T (scipy qhull._barycentric_coordinates)
__device__ __forceinline__ void barycentric(const Lattice &L, const Tri &T, double X, double Y, double b[3]) {
    const double x2 = L.x[T.v[2]], y2 = L.y[T.v[2]];
    const double a00 = L.x[T.v[0]] - x2, a01 = L.x[T.v[1]] - x2, a10 = L.y[T.v[0]] - y2, a11 = L.y[T.v[1]] - y2;
    const double det = a00 * a11 - a01 * a10, dx = X - x2, dy = Y - y2;
    b[0] = (dx * a11 - a01 * dy) / det;
    b[1] = (a00 * dy - dx * a10) / det;
    b[2] = 1.0 - b[0] - b[1];
}
// the triangle on the other side of the edge opposite vertex kk of T (false: hull edge)
__device__ bool neighbour_tri(const Lattice &L, const Tri &T, int kk, Tri &N) {
    const int u = T.k[(kk + 1) % 3], v = T.k[(kk + 2) % 3];  // corners of the shared edge
    const int lo = min(u, v), hi = max(u, v);
    if (hi - lo == 2) {  // a diagonal of the quad: the quad's other triangle
        return L.flag(T.r, T.c) <= 2 && load_tri(L, T.r, T.c, T.t ^ 1, N);
    }
    // a lattice side: {0,1} top -> quad (r-1, c) side {3,2}; {1,2} right -> (r, c+1) side {0,3};
    // {2,3} bottom -> (r+1, c) side {1,0}; {0,3} left -> (r, c-1) side {1,2}
    int r2 = T.r, c2 = T.c, a, b;
    if (lo == 0 && hi == 1) { r2 -= 1; a = 3; b = 2; }
    else if (lo == 1 && hi == 2) { c2 += 1; a = 0; b = 3; }
    else if (lo == 2 && hi == 3) { r2 += 1; a = 1; b = 0; }
    else { c2 -= 1; a = 1; b = 2; }
    for (int t = 0; t < 2; ++t) {
        if (!load_tri(L, r2, c2, t, N)) continue;
        bool ha = false, hb = false;
        for (int q = 0; q < 3; ++q) {
            ha |= N.k[q] == a;
            hb |= N.k[q] == b;
        }
        if (ha && hb) return true;
    }
    return false;
}

// scipy interpnd.pyx _clough_tocher_2d_single for two fields at once (shared geometry)
__device__ void clough_tocher(const Lattice &L, const Tri &T, const double b[3], const double *fa,
                              const double *fp, const double *grad, double &wa, double &wp) {
    const double p0x = L.x[T.v[0]], p0y = L.y[T.v[0]], p1x = L.x[T.v[1]], p1y = L.y[T.v[1]],
                 p2x = L.x[T.v[2]], p2y = L.y[T.v[2]];
    const double e12x = p1x - p0x, e12y = p1y - p0y, e23x = p2x - p1x, e23y = p2y - p1y,
                 e31x = p0x - p2x, e31y = p0y - p2y;
    double g[3];
    for (int k = 0; k < 3; ++k) {
        Tri N;
        if (!neighbour_tri(L, T, k, N)) {
            g[k] = -0.5;
            continue;
        }
        const double yx = (L.x[N.v[0]] + L.x[N.v[1]] + L.x[N.v[2]]) / 3, yy = (L.y[N.v[0]] + L.y[N.v[1]] + L.y[N.v[2]]) / 3;
        double c[3];
        barycentric(L, T, yx, yy, c);
        if (k == 0) g[k] = (2 * c[2] + c[1] - 1) / (2 - 3 * c[2] - 3 * c[1]);
        else if (k == 1) g[k] = (2 * c[0] + c[2] - 1) / (2 - 3 * c[0] - 3 * c[2]);
        else g[k] = (2 * c[1] + c[0] - 1) / (2 - 3 * c[1] - 3 * c[0]);
    }
    double minval = fmin(b[0], fmin(b[1], b[2]));
    const double b1 = b[0] - minval, b2 = b[1] - minval, b3 = b[2] - minval, b4 = 3 * minval;
    for (int fld = 0; fld < 2; ++fld) {
        const double *f = fld ? fp : fa;
        const double f1 = f[T.v[0]], f2 = f[T.v[1]], f3 = f[T.v[2]];
        const double *d0 = grad + 4ll * T.v[0] + 2 * fld, *d1 = grad + 4ll * T.v[1] + 2 * fld,
                     *d2 = grad + 4ll * T.v[2] + 2 * fld;
        const double df12 = +(d0[0] * e12x + d0[1] * e12y), df21 = -(d1[0] * e12x + d1[1] * e12y);
        const double df23 = +(d1[0] * e23x + d1[1] * e23y), df32 = -(d2[0] * e23x + d2[1] * e23y);
        const double df31 = +(d2[0] * e31x + d2[1] * e31y), df13 = -(d0[0] * e31x + d0[1] * e31y);
        const double c3000 = f1, c2100 = (df12 + 3 * c3000) / 3, c2010 = (df13 + 3 * c3000) / 3;
        const double c0300 = f2, c1200 = (df21 + 3 * c0300) / 3, c0210 = (df23 + 3 * c0300) / 3;
        const double c0030 = f3, c1020 = (df31 + 3 * c0030) / 3, c0120 = (df32 + 3 * c0030) / 3;
        const double c2001 = (c2100 + c2010 + c3000) / 3, c0201 = (c1200 + c0300 + c0210) / 3,
                     c0021 = (c1020 + c0120 + c0030) / 3;
        const double c0111 = (g[0] * (-c0300 + 3 * c0210 - 3 * c0120 + c0030) +
                              (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2;
        const double c1011 = (g[1] * (-c0030 + 3 * c1020 - 3 * c2010 + c3000) +
                              (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2;
        const double c1101 = (g[2] * (-c3000 + 3 * c2100 - 3 * c1200 + c0300) +
                              (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2;
        const double c1002 = (c1101 + c1011 + c2001) / 3, c0102 = (c1101 + c0111 + c0201) / 3,
                     c0012 = (c1011 + c0111 + c0021) / 3;
        const double c0003 = (c1002 + c0102 + c0012) / 3;
        const double w =
            b1 * b1 * b1 * c3000 + 3 * b1 * b1 * b2 * c2100 + 3 * b1 * b1 * b3 * c2010 +
            3 * b1 * b1 * b4 * c2001 + 3 * b1 * b2 * b2 * c1200 + 6 * b1 * b2 * b4 * c1101 +
            3 * b1 * b3 * b3 * c1020 + 6 * b1 * b3 * b4 * c1011 + 3 * b1 * b4 * b4 * c1002 +
            b2 * b2 * b2 * c0300 + 3 * b2 * b2 * b3 * c0210 + 3 * b2 * b2 * b4 * c0201 +
            3 * b2 * b3 * b3 * c0120 + 6 * b2 * b3 * b4 * c0111 + 3 * b2 * b4 * b4 * c0102 +
            b3 * b3 * b3 * c0030 + 3 * b3 * b3 * b4 * c0021 + 3 * b3 * b4 * b4 * c0012 +
            b4 * b4 * b4 * c0003;
        if (fld) wp = w; else wa = w;
    }
}

// does quad (r, c) contain (X, Y)?  Tries its triangles; on success T and b are set.
__device__ bool locate_in_quad(const Lattice &L, int r, int c, double X, double Y, Tri &T, double b[3]) {
    const double eps = 1e-12;
    for (int t = 0; t < 2; ++t) {
        Tri C;
        if (!load_tri(L, r, c, t, C)) continue;
        double bb[3];
        barycentric(L, C, X, Y, bb);
        if (bb[0] >= -eps && bb[1] >= -eps && bb[2] >= -eps) {
            T = C;
            b[0] = bb[0]; b[1] = bb[1]; b[2] = bb[2];
            return true;
        }
    }
    return false;
}

// One output grid point per thread: out[i][j] at (gx[i], gy[j]) (np.mgrid ordering: x along axis
// 0).  Point location is a Newton iteration on the bilinear map of the lattice quads started at
// the node nearest to the beam centre (the map is smooth and monotone), confirmed by the
// barycentric test in the 3 x 3 quads around the result; a full scan of the quads is the
// fallback.  Outside the triangulation -> NaN -> 0, like griddata's fill value after
// ray_trace.py:1698 / :1707; r >= r_zero from the beam centre -> 0 (:1692-1697).
__global__ void g2d_eval_kernel(Lattice L, const double *__restrict__ fa, const double *__restrict__ fp,
                                const double *__restrict__ grad, const double *__restrict__ sc,
                                const double *__restrict__ gx, const double *__restrict__ gy, int m,
                                double r_zero2, double *__restrict__ out_a, double *__restrict__ out_p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * m) return;
    const int i = idx / m, j = idx % m;
    const double X = gx[i], Y = gy[j];
    const double cx = sc[SC_CX], cy = sc[SC_CY];
    double va = 0.0, vp = 0.0;
    if (!((X - cx) * (X - cx) + (Y - cy) * (Y - cy) >= r_zero2)) {
        const int n = L.n;
        const long long centre = (long long)sc[SC_CENTRE];
        double u = (double)(centre / n), v = (double)(centre % n);  // continuous (row, column)
        Tri T;
        double b[3];
        bool found = false;
        for (int it = 0; it < 64 && !found; ++it) {
            int r = (int)floor(u), c = (int)floor(v);
            r = max(0, min(n - 2, r));
            c = max(0, min(n - 2, c));
            const int v0 = L.corner(r, c, 0), v1 = L.corner(r, c, 1), v2 = L.corner(r, c, 2), v3 = L.corner(r, c, 3);
            if (v0 < 0 || v1 < 0 || v2 < 0 || v3 < 0) break;  // walked off the selected region
            const double s = u - r, t = v - c;  // s along rows (q0 -> q3), t along columns (q0 -> q1)
            const double x0 = L.x[v0], y0 = L.y[v0], x1 = L.x[v1], y1 = L.y[v1], x2 = L.x[v2], y2 = L.y[v2],
                         x3 = L.x[v3], y3 = L.y[v3];
            const double Px = (1 - s) * ((1 - t) * x0 + t * x1) + s * ((1 - t) * x3 + t * x2);
            const double Py = (1 - s) * ((1 - t) * y0 + t * y1) + s * ((1 - t) * y3 + t * y2);
            const double dxs = (1 - t) * (x3 - x0) + t * (x2 - x1), dys = (1 - t) * (y3 - y0) + t * (y2 - y1);
            const double dxt = (1 - s) * (x1 - x0) + s * (x2 - x3), dyt = (1 - s) * (y1 - y0) + s * (y2 - y3);
            const double det = dxs * dyt - dxt * dys, rx = X - Px, ry = Y - Py;
            const double ds = (rx * dyt - dxt * ry) / det, dt = (dxs * ry - rx * dys) / det;
            u += ds;
            v += dt;
            if (fabs(ds) < 0.25 && fabs(dt) < 0.25) {
                const int rc = (int)floor(u), cc = (int)floor(v);
                for (int q = 0; q < 9 && !found; ++q) {  // the quad of the solution first, then its ring
                    const int order[9] = {4, 1, 3, 5, 7, 0, 2, 6, 8};
                    found = locate_in_quad(L, rc + order[q] / 3 - 1, cc + order[q] % 3 - 1, X, Y, T, b);
                }
            }
        }
        if (!found) {
            for (int r = 0; r < n - 1 && !found; ++r)
                for (int c = 0; c < n - 1 && !found; ++c) found = locate_in_quad(L, r, c, X, Y, T, b);
        }
        if (found) {
            clough_tocher(L, T, b, fa, fp, grad, va, vp);
            if (va != va) va = 0.0;
            if (vp != vp) vp = 0.0;
        }
    }
    out_a[idx] = va;
    out_p[idx] = vp;
}

// ---- sim_data ---------------------------------------------------------------------------------
// z_amp = |a|^T / max|a|, z_ph = arctan2(Im, Re) of (a exp(i p))^T  (ray_trace.py:1725-1740)
__global__ void sd_max_kernel(const double *__restrict__ a, int n2, double *mx) {
    double best = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
        best = fmax(best, fabs(a[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomic_max_double(mx, best);
}
__global__ void sd_prep_kernel(const double *__restrict__ a, const double *__restrict__ p, int n,
                               const double *__restrict__ mx, double *__restrict__ z_amp,
                               double *__restrict__ z_ph) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int iy = idx / n, ix = idx % n;  // output [iy][ix] = input [ix][iy]
    const double av = a[(long long)ix * n + iy], pv = p[(long long)ix * n + iy];
    double sn, cs;
    sincos(pv, &sn, &cs);
    z_amp[idx] = fabs(av) / mx[0];
    z_ph[idx] = atan2(av * sn, av * cs);
}
// C[m][q] = sum_k A[m][k] B(k, q), B given as [k][q] (TRANS_B = false) or [q][k] (true); fp64,
// one output per thread: the operands are a few hundred KB and live in L1/L2
template <bool TRANS_B>
__global__ void sd_gemm_kernel(const double *__restrict__ A, const double *__restrict__ B, int M, int K,
                               int Q, double *__restrict__ C) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x, mrow = blockIdx.y;
    if (q >= Q) return;
    double acc = 0.0;
    for (int k = 0; k < K; ++k)
        acc = fma(A[(long long)mrow * K + k], TRANS_B ? B[(long long)q * K + k] : B[(long long)k * Q + q], acc);
    C[(long long)mrow * Q + q] = acc;
}
__global__ void sd_signed_max_kernel(const double *__restrict__ a, int n2, double *mx) {
    double best = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
        best = fmax(best, a[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomic_max_double(mx, best);
}
__global__ void sd_scale_kernel(double *a, int n2, const double *mx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) a[i] = a[i] / mx[0];
}
__global__ void sd_set_kernel(double *dst, double v0, double v1) {
    dst[0] = v0;
    dst[1] = v1;
}

}  // namespace sosat

using namespace sosat;

extern "C" size_t sosat_sim2d_workspace_bytes(int64_t n_pts, int n_scan) {
    if (n_pts < 0 || n_scan < 2) return 0;
    const size_t nodes = (size_t)n_scan * n_scan, quads = (size_t)(n_scan - 1) * (n_scan - 1);
    return 256 + ((nodes * 4 + 255) / 256) * 256 + ((quads + 255) / 256) * 256 + (size_t)n_pts * 32;
}

namespace {
struct Sim2dWs {
    double *sc, *grad;
    int *node;
    unsigned char *quad;
};
Sim2dWs carve_sim2d(void *ws, int64_t n_pts, int n_scan) {
    char *p = (char *)ws;
    Sim2dWs w;
    const size_t nodes = (size_t)n_scan * n_scan, quads = (size_t)(n_scan - 1) * (n_scan - 1);
    w.sc = (double *)p;
    p += 256;
    w.node = (int *)p;
    p += ((nodes * 4 + 255) / 256) * 256;
    w.quad = (unsigned char *)p;
    p += ((quads + 255) / 256) * 256;
    w.grad = (double *)p;
    return w;
}
int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace

extern "C" int sosat_sim2d_prepare(const double *d_x, const double *d_y, const double *d_a,
                                   const double *d_p, const int64_t *d_ray_index, int64_t n_pts,
                                   int n_scan, double r_select, int sweeps, void *d_workspace,
                                   size_t workspace_bytes, double *d_scalars, void *stream) {
    SOSAT_REQUIRE(d_x && d_y && d_a && d_p && d_ray_index && d_workspace && d_scalars,
                  "sosat_sim2d_prepare: null pointer");
    SOSAT_REQUIRE(n_pts >= 1 && n_scan >= 2 && n_scan <= 46340, "n_pts / n_scan out of range");
    SOSAT_REQUIRE(sweeps >= 1 && sweeps <= 10000, "sweeps out of range");
    SOSAT_REQUIRE(workspace_bytes >= sosat_sim2d_workspace_bytes(n_pts, n_scan) &&
                      ((uintptr_t)d_workspace & 255) == 0,
                  "sim2d workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Sim2dWs w = carve_sim2d(d_workspace, n_pts, n_scan);
    const long long nodes = (long long)n_scan * n_scan, quads = (long long)(n_scan - 1) * (n_scan - 1);
    g2d_init_kernel<<<blocks_for(nodes), 256, 0, st>>>(w.sc, w.node, nodes);
    g2d_sum_kernel<<<blocks_for(n_pts), 256, 0, st>>>(d_x, d_y, n_pts, w.sc);
    g2d_select_kernel<<<blocks_for(n_pts), 256, 0, st>>>(d_x, d_y, (const long long *)d_ray_index, n_pts,
                                                         r_select * r_select, w.sc, w.node);
    g2d_centre_kernel<<<blocks_for(n_pts), 256, 0, st>>>(d_x, d_y, (const long long *)d_ray_index, n_pts,
                                                         r_select * r_select, w.sc);
    g2d_quad_kernel<<<blocks_for(quads), 256, 0, st>>>(d_x, d_y, w.node, n_scan, w.quad);
    SOSAT_CUDA_OK(cudaMemsetAsync(w.grad, 0, (size_t)n_pts * 32, st));
    Lattice L{d_x, d_y, w.node, w.quad, n_scan};
    for (int s = 0; s < sweeps; ++s)
        for (int colour = 0; colour < 4; ++colour)
            g2d_gradient_kernel<<<blocks_for(nodes / 4 + 1), 256, 0, st>>>(L, d_a, d_p, w.grad, colour);
    SOSAT_CUDA_OK(cudaMemcpyAsync(d_scalars, w.sc, 8 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_sim2d_eval(const double *d_x, const double *d_y, const double *d_a,
                                const double *d_p, int64_t n_pts, int n_scan, const void *d_workspace,
                                const double *d_gx, const double *d_gy, int m, double r_zero,
                                double *d_out_a, double *d_out_p, void *stream) {
    SOSAT_REQUIRE(d_x && d_y && d_a && d_p && d_workspace && d_gx && d_gy && d_out_a && d_out_p,
                  "sosat_sim2d_eval: null pointer");
    SOSAT_REQUIRE(n_pts >= 1 && n_scan >= 2 && m >= 1 && m <= 16384, "size out of range");
    Sim2dWs w = carve_sim2d(const_cast<void *>(d_workspace), n_pts, n_scan);
    Lattice L{d_x, d_y, w.node, w.quad, n_scan};
    g2d_eval_kernel<<<(m * m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        L, d_a, d_p, w.grad, w.sc, d_gx, d_gy, m, r_zero * r_zero, d_out_a, d_out_p);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" int sosat_sim_data(const double *d_a, const double *d_p, int n, const double *d_Mx,
                              const double *d_My, int m, double *d_a_data, double *d_p_data,
                              double *d_work, void *stream) {
    SOSAT_REQUIRE(d_a && d_p && d_Mx && d_My && d_a_data && d_p_data && d_work,
                  "sosat_sim_data: null pointer");
    SOSAT_REQUIRE(n >= 4 && n <= 8192 && m >= 1 && m <= 16384, "size out of range");
    cudaStream_t st = (cudaStream_t)stream;
    // d_work: [2 scalars | z_amp n*n | z_ph n*n | t1 m*n]
    double *mx = d_work, *z_amp = d_work + 32, *z_ph = z_amp + (size_t)n * n, *t1 = z_ph + (size_t)n * n;
    sd_set_kernel<<<1, 1, 0, st>>>(mx, 0.0, -INFINITY);
    sd_max_kernel<<<blocks_for((long long)n * n), 256, 0, st>>>(d_a, n * n, mx);
    sd_prep_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(d_a, d_p, n, mx, z_amp, z_ph);
    const dim3 g1((n + 127) / 128, m), g2((m + 127) / 128, m);
    // R[ky][kx] = sum_{iy, ix} My[ky][iy] z[iy][ix] Mx[kx][ix]
    sd_gemm_kernel<false><<<g1, 128, 0, st>>>(d_My, z_amp, m, n, n, t1);
    sd_gemm_kernel<true><<<g2, 128, 0, st>>>(t1, d_Mx, m, n, m, d_a_data);
    sd_gemm_kernel<false><<<g1, 128, 0, st>>>(d_My, z_ph, m, n, n, t1);
    sd_gemm_kernel<true><<<g2, 128, 0, st>>>(t1, d_Mx, m, n, m, d_p_data);
    sd_signed_max_kernel<<<blocks_for((long long)m * m), 256, 0, st>>>(d_a_data, m * m, mx + 1);
    sd_scale_kernel<<<(m * m + 255) / 256, 256, 0, st>>>(d_a_data, m * m, mx + 1);
    SOSAT_CUDA_OK(cudaGetLastError());
    return SOSAT_OK;
}

extern "C" size_t sosat_sim_data_work_doubles(int n, int m) {
    return 32 + 2 * (size_t)n * n + (size_t)m * n;
}
