"""Drop-in for the hot-path functions of the reference's ``sosat_optics/ray_trace.py``.

Same names, arguments and return values as the reference; the arithmetic runs in
libsosat_b200.so on the GPU (csrc/trace.cu).  PyTorch is used only to own device memory and
the CUDA stream.

======================  ==========================  ==========================================
function                reference                   device code
======================  ==========================  ==========================================
``rx_to_lyot``          ray_trace.py:1070-1591      ``sosat_trace_rays``  (trace_rows_kernel)
``getNearField``        ray_trace.py:1594-1614      ``sosat_near_field_arrays``
``near_field_binned``   -- (north_star, row a7)     ``sosat_trace_bin``   (trace_bin_kernel)
``get_fwhm`` etc.       ray_trace.py:1648-1672,     host numpy helpers over returned arrays (the
``ff_fwhm_est``         1757-1779                   reference's signatures) ...
``near_field_samples``  -- (north_star, irregular)    ``sosat_trace_samples``: float4 {x, y, Re b, Im b}
                        route)                      per ray, the input of the direct-sum far field
``near_field_metrics``  same reductions, on device  ``sosat_near_field_metrics`` (reduce.cu): one
                        for a whole frequency sweep trace, scalars per frequency come back
``ff_fwhm_device``      ff_fwhm_est on a CUDA field ``sosat_ff_fwhm_indices`` (reduce.cu)
``sim2d``               ray_trace.py:1675-1717      ``sosat_sim2d_prepare`` / ``sosat_sim2d_eval``
                                                    (gridder.cu): scipy's cubic griddata restated
                                                    on the ray lattice
``sim_data``            ray_trace.py:1720-1754      ``sosat_sim_data``: the bicubic spline resampling
                                                    as two fp64 contractions
======================  ==========================  ==========================================

``plot`` / ``col`` arguments are accepted and ignored (plotting is out of scope).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib, geometry

__all__ = ["rx_to_lyot", "getNearField", "get_sim", "get_beam_cent", "get_fwhm",
           "get_angle_out", "ff_fwhm_est", "near_field_binned", "trace_stats",
           "near_field_metrics", "near_field_samples", "ff_fwhm_device", "ff_fwhm_indices_device",
           "beam_sweep", "sim2d",
           "sim_data",
           "SimOutput", "BinnedField"]


_torch_ok = None


def _torch():
    global _torch_ok
    if _torch_ok is not None:   # a device that was there stays there: check once per process
        return _torch_ok
    import torch
    if not torch.cuda.is_available():
        raise _lib.SosatError("no CUDA device visible: sosat_optics_b200 runs only on the GPU "
                              "(there is deliberately no CPU fallback)")
    _torch_ok = torch
    return torch


def _stream(torch) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_geom_cache = {}


def _upload_geometry(lib, torch, tele_geo) -> None:
    """sosat_set_geometry, skipped when this device already holds identical constants."""
    g = geometry.make_geom(tele_geo)
    dev = torch.cuda.current_device()
    key = bytes(g)
    if _geom_cache.get(dev) != key:
        _lib.check(lib.sosat_set_device(dev))
        _lib.check(lib.sosat_set_geometry(ctypes.byref(g), _stream(torch)))
        _geom_cache[dev] = key


def _rx_array(P_rx) -> np.ndarray:
    rx = np.ascontiguousarray(np.asarray(P_rx, dtype=np.float64).reshape(-1))
    if rx.size != 3:
        raise ValueError("receiver position must have 3 components [x, y, z] in mm")
    return rx


def _trace_device(P_rx, tele_geo, ray_begin=0, ray_end=None):
    """Run the tracer; returns (out [17, n] cuda float64 tensor, stats host ndarray)."""
    lib = _lib.require_gpu()
    torch = _torch()
    n_scan = int(tele_geo.n_scan)
    total = n_scan * n_scan
    if ray_end is None:
        ray_end = total
    n = ray_end - ray_begin
    _upload_geometry(lib, torch, tele_geo)
    freq = geometry.make_freq(tele_geo)
    rx = _rx_array(P_rx)
    out = torch.empty((_lib.OUT_ROWS, max(n, 0)), dtype=torch.float64, device="cuda")
    stats = torch.zeros(_lib.NUM_STATS, dtype=torch.float64, device="cuda")
    _lib.check(lib.sosat_trace_rays(
        rx.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(freq), n_scan,
        ray_begin, ray_end, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stats.data_ptr()),
        _stream(torch)))
    h_stats = _lib.to_numpy(stats)
    if h_stats[_lib.STAT_N_BRACKET_FAIL] > 0:
        # scipy.optimize.brentq's error, which aborts the reference's whole run
        raise ValueError("f(a) and f(b) must have different signs")
    return out, h_stats


def rx_to_lyot(P_rx, tele_geo, plot=False, col="b"):
    """Trace ``n_scan**2`` rays from the feed at ``P_rx`` [mm] through the three lenses and the
    Lyot stop to the source plane (ray_trace.py:1070).  Returns ``out[17, n_scan**2]`` float64:
    rows 0-2 position on the source plane, 3 optical path length, 4 amplitude, 5-7 last surface
    normal, 8-10 exit direction, 11-16 zero."""
    out, _ = _trace_device(P_rx, tele_geo)
    return _lib.to_numpy(out)


class SimOutput:
    """Same attribute bag as the class defined inside the reference's getNearField
    (ray_trace.py:1606-1612)."""

    def __init__(self, x_sim, y_sim, a_sim, p_sim, tan_og):
        self.a_sim = a_sim
        self.p_sim = p_sim
        self.x_sim = x_sim
        self.y_sim = y_sim
        self.tan_og = tan_og


def getNearField(tele_geo, rx, plot=False):
    """Get the near field of a receiver feed (ray_trace.py:1594-1614): positions [cm],
    amplitude and mean-removed wrapped phase of the valid rays, and the mean exit direction."""
    lib = _lib.require_gpu()
    torch = _torch()
    out, _ = _trace_device(rx, tele_geo)
    n = out.shape[1]
    xyap = torch.empty((4, n), dtype=torch.float64, device="cuda")
    mask = torch.empty(n, dtype=torch.uint8, device="cuda")
    tan_og = torch.empty(3, dtype=torch.float64, device="cuda")
    work = torch.empty(_lib.NUM_STATS, dtype=torch.float64, device="cuda")
    _lib.check(lib.sosat_near_field_arrays(
        ctypes.c_void_p(out.data_ptr()), n, geometry.scalar(tele_geo.k),
        ctypes.c_void_p(xyap.data_ptr()), ctypes.c_void_p(mask.data_ptr()),
        ctypes.c_void_p(tan_og.data_ptr()), ctypes.c_void_p(work.data_ptr()), _stream(torch)))
    keep = mask.bool()
    sel = _lib.to_numpy(xyap[:, keep])  # xx = np.where(out[4] != 0): order-preserving
    sb = SimOutput(sel[0].copy(), sel[1].copy(), sel[2].copy(), sel[3].copy(),
                   _lib.to_numpy(tan_og))
    # where each returned ray sits in the n_scan x n_scan launch lattice (ray ii = iphi * n_scan +
    # itheta, ray_trace.py:1086-1091): what sim2d needs to triangulate the rays without Qhull
    sb.ray_index = _lib.to_numpy(torch.nonzero(keep).reshape(-1))
    sb.n_scan = int(tele_geo.n_scan)
    return sb


# --- host helpers over the returned arrays (ray_trace.py:1648-1672, 1757-1779) ------------
def get_sim(sb):
    return sb.a_sim, sb.p_sim, sb.x_sim, sb.y_sim


def get_beam_cent(sb):
    return [np.mean(sb.x_sim), np.mean(sb.y_sim)]


def get_fwhm(sb):
    """Half-extent of the rays above half power; note the reference's swapped x/y naming."""
    a_sim, _, x_sim, y_sim = get_sim(sb)
    beam_cent = get_beam_cent(sb)
    above = np.where(a_sim ** 2 > np.max(a_sim ** 2) / 2)
    fwhm_nf_x = abs(y_sim - beam_cent[1])[above].max()
    fwhm_nf_y = abs(x_sim - beam_cent[0])[above].max()
    return fwhm_nf_x, fwhm_nf_y


def get_angle_out(sb):
    theta_x = abs(90 - np.rad2deg(np.arctan2(sb.tan_og[1], sb.tan_og[2])))
    theta_y = abs(90 - np.rad2deg(np.arctan2(sb.tan_og[1], sb.tan_og[0])))
    return [theta_x, theta_y]


def ff_fwhm_est(beam_fft, x_ang, y_ang):
    """FWHM [arcmin] of |B|^2 through the peak row and column (ray_trace.py:1757-1779)."""
    a = abs(beam_fft) ** 2 / np.max(abs(beam_fft) ** 2)
    x_out, y_out = y_ang, x_ang
    indx = np.where(abs(a) == np.max(abs(a)))
    x = y_out[indx[0][0], :] * 60 * 180 / np.pi
    y = abs(a)[indx[0][0], :] / np.max(abs(a))
    fwhm1 = abs(x[np.where(y > 0.5)][0] - x[np.where(y > 0.5)][-1])
    x = x_out[:, indx[1][0]] * 60 * 180 / np.pi
    y = abs(a)[:, indx[1][0]] / np.max(abs(a))
    fwhm2 = abs(x[np.where(y > 0.5)][0] - x[np.where(y > 0.5)][-1])
    return (fwhm1 + fwhm2) / 2


# --- compact per-ray samples for the irregular-sample far field (north_star) -------------------
def near_field_samples(tele_geo, rx, *, device: bool = False):
    """Trace the bundle of one feed and return the near field ON THE RAY POSITIONS as samples
    ``[n_scan**2, 4]`` float32 ``{x_sim [cm], y_sim [cm], Re b, Im b}`` with ``b = a exp(i p)``,
    ``p = k (OPL - <OPL>)/1e3/2`` (invalid rays are all-zero rows, so they drop out of any sum),
    plus the statistics row.  This is the input format of :func:`opt_analyze.far_field_irregular`
    (pass the tensor with ``device=True`` to stay on the GPU): no gridding between the ray trace
    and the far field.  One trace: the phase is first referred to OPL = 0 inside the kernel and
    the global factor ``exp(-i k <OPL>/2e3)`` is applied afterwards."""
    lib = _lib.require_gpu()
    torch = _torch()
    n_scan = int(tele_geo.n_scan)
    n = n_scan * n_scan
    _upload_geometry(lib, torch, tele_geo)
    freq = geometry.make_freq(tele_geo)
    rxa = _rx_array(rx)
    xyb = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    stats = torch.zeros(_lib.NUM_STATS, dtype=torch.float64, device="cuda")
    _lib.check(lib.sosat_trace_samples(
        rxa.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(freq), n_scan, 0, n, 0.0,
        ctypes.c_void_p(xyb.data_ptr()), ctypes.c_void_p(stats.data_ptr()), _stream(torch)))
    h_stats = _lib.to_numpy(stats)
    if h_stats[_lib.STAT_N_BRACKET_FAIL] > 0:
        raise ValueError("f(a) and f(b) must have different signs")
    nv = h_stats[_lib.STAT_N_VALID]
    mean_opl = h_stats[_lib.STAT_SUM_OPL] / nv if nv > 0 else 0.0
    phi0 = float(np.mod(geometry.scalar(tele_geo.k) / 1e3 / 2 * mean_opl, 2 * np.pi))
    b = torch.view_as_complex(xyb[:, 2:4].contiguous()) * complex(np.cos(phi0), -np.sin(phi0))
    xyb[:, 2:4] = torch.view_as_real(b)
    return (xyb if device else _lib.to_numpy(xyb)), h_stats


# --- on-device reductions for sweeps -------------------------------
def near_field_metrics(tele_geo, rx, freqs: Optional[Sequence] = None):
    """``get_beam_cent`` / ``get_fwhm`` / ``get_angle_out`` of ``getNearField(tele_geo, rx)``
    evaluated on the device for a whole list of frequencies with ONE trace (the ray geometry is
    frequency independent): the reference's sat_reqs sweep re-traces every frequency.

    ``freqs``: ``None`` -> the frequency set on ``tele_geo``; else a sequence of
    ``(k, th_fwhp_x, th_fwhp_y)``.  Returns a dict: ``n_valid``, ``beam_cent`` [cm] (2),
    ``tan_og`` (3), ``angle_out`` [deg] (2), ``fwhm`` [cm] ``[F, 2]`` in the order
    ``get_fwhm`` returns (its swapped x/y naming kept)."""
    lib = _lib.require_gpu()
    torch = _torch()
    out, _ = _trace_device(rx, tele_geo)
    specs = list(freqs) if freqs is not None else [
        (tele_geo.k, tele_geo.th_fwhp_x, tele_geo.th_fwhp_y)]
    n_scan = int(tele_geo.n_scan)
    res_all = []
    head = None
    for f0 in range(0, len(specs), _lib.MAX_FREQS):
        chunk = geometry.make_freqs(specs[f0:f0 + _lib.MAX_FREQS])
        result = torch.empty(6 + 2 * _lib.MAX_FREQS, dtype=torch.float64, device="cuda")
        work = torch.empty(8 + _lib.MAX_FREQS, dtype=torch.float64, device="cuda")
        _lib.check(lib.sosat_near_field_metrics(
            ctypes.c_void_p(out.data_ptr()), out.shape[1], n_scan, 0, geometry.ot_geo.CONE_HALF_ANGLE,
            chunk, len(chunk), ctypes.c_void_p(result.data_ptr()), ctypes.c_void_p(work.data_ptr()),
            _stream(torch)))
        h = _lib.to_numpy(result)
        head = h[:6]
        res_all.append(h[6:6 + 2 * len(chunk)].reshape(-1, 2))
    tan_og = head[3:6].copy()
    sb = SimOutput(None, None, None, None, tan_og)
    return {"n_valid": int(head[0]), "beam_cent": head[1:3].copy(), "tan_og": tan_og,
            "angle_out": np.array(get_angle_out(sb)), "fwhm": np.concatenate(res_all, axis=0)}


def ff_fwhm_indices_device(B):
    """The searches of ``ff_fwhm_est`` for a stack of far fields ``B`` [K, N, N] (CUDA complex64)
    in one launch per search and ONE read-back: returns ``(idx int32 [K, 6], peak_abs2 float32
    [K])`` as numpy arrays, ``idx`` = peak row, peak column, first / last column above half power
    in the peak row, first / last row above half power in the peak column.  Raises ``ValueError``
    if a field has no finite non-zero cell (NaN in the near field)."""
    lib = _lib.require_gpu()
    torch = _torch()
    k, n = int(B.shape[0]), int(B.shape[-1])
    res = torch.empty((k, 8), dtype=torch.int32, device="cuda")  # [idx x 6 | val bits | pad]
    idx = torch.empty((k, 6), dtype=torch.int32, device="cuda")
    val = torch.empty(k, dtype=torch.float32, device="cuda")
    work = torch.empty(k, dtype=torch.int64, device="cuda")
    _lib.check(lib.sosat_ff_fwhm_indices_batch(
        ctypes.c_void_p(B.data_ptr()), n, k, ctypes.c_void_p(idx.data_ptr()),
        ctypes.c_void_p(val.data_ptr()), ctypes.c_void_p(work.data_ptr()), _stream(torch)))
    res[:, :6] = idx
    res[:, 6] = val.view(torch.int32)
    h = _lib.to_numpy(res)
    h_idx, h_val = h[:, :6].copy(), h[:, 6].copy().view(np.float32)
    bad = np.nonzero((h_idx < 0).any(axis=1))[0]
    if bad.size:
        raise ValueError(f"far field {bad.tolist()[:8]} has no finite non-zero cell (NaN or all-zero "
                         "near field): no beam peak to measure")
    return h_idx, h_val


def ff_fwhm_device(beam_fft, x_ang, y_ang):
    """``ff_fwhm_est`` (ray_trace.py:1757-1779) for a far field that lives on the GPU
    (``far_field(..., device=True)``): the peak and half-power searches run on the device and
    only six indices come back.  ``x_ang`` / ``y_ang``: the label grids of
    ``coords_spat_to_ang`` (host arrays)."""
    torch = _torch()
    B = beam_fft if not isinstance(beam_fft, np.ndarray) else torch.from_numpy(
        np.ascontiguousarray(beam_fft.astype(np.complex64))).to("cuda")
    B = B.to(device="cuda", dtype=torch.complex64).contiguous()
    idx, _ = ff_fwhm_indices_device(B.unsqueeze(0))
    pr, pc, c0, c1, r0, r1 = (int(v) for v in idx[0])
    x_out, y_out = y_ang, x_ang  # the reference's swap, ray_trace.py:1759-1760
    scale = 60 * 180 / np.pi
    fwhm1 = abs(y_out[pr, c0] * scale - y_out[pr, c1] * scale)
    fwhm2 = abs(x_out[r0, pc] * scale - x_out[r1, pc] * scale)
    return (fwhm1 + fwhm2) / 2


# --- the notebooks' scattered -> grid glue, on the device -------------
class _Grid2d:
    def __init__(self, x_sim, y_sim, a_sim, p_sim):
        self.a_sim, self.p_sim, self.x_sim, self.y_sim = a_sim, p_sim, x_sim, y_sim


SIM2D_GRID = 100          # np.mgrid[...:100j], ray_trace.py:1686-1689
SIM2D_R_SELECT = 21.0     # cm, :1684
SIM2D_R_ZERO = 20.0       # cm, :1693
SIM2D_SWEEPS = 40         # four-colour Gauss-Seidel sweeps (scipy stops at a 1e-6 change: 10-13)


def sim2d(sb):
    """Cubic ``griddata`` of amplitude and phase on a 100 x 100 grid spanning the valid rays within
    21 cm of the beam centre, zero beyond 20 cm (ray_trace.py:1675-1717), on the GPU.

    ``scipy.interpolate.griddata(method="cubic")`` is Clough-Tocher interpolation on the Delaunay
    triangulation of the points with globally estimated gradients; ``csrc/gridder.cu`` restates
    that algorithm for ray positions that form a warped structured lattice, which is why ``sb``
    must come from this package's :func:`getNearField` (it carries ``ray_index`` / ``n_scan``,
    the place of every ray in the launch lattice).  Returns the reference's attribute bag
    (``x_sim, y_sim`` = the ``np.mgrid`` coordinates, ``a_sim, p_sim`` [100, 100])."""
    lib = _lib.require_gpu()
    torch = _torch()
    idx = getattr(sb, "ray_index", None)
    n_scan = getattr(sb, "n_scan", None)
    if idx is None or n_scan is None:
        raise ValueError("sim2d needs the SimOutput of sosat_optics_b200.ray_trace.getNearField "
                         "(its ray_index / n_scan attributes place the rays in the launch lattice)")
    n = len(sb.a_sim)
    if not (len(sb.x_sim) == len(sb.y_sim) == len(sb.p_sim) == len(idx) == n) or n == 0:
        raise ValueError("sim2d: x_sim, y_sim, a_sim, p_sim and ray_index must have one entry per ray")
    idx = np.asarray(idx, dtype=np.int64)
    if int(n_scan) < 2 or idx.min() < 0 or idx.max() >= int(n_scan) ** 2 or np.any(np.diff(idx) <= 0):
        raise ValueError("sim2d: ray_index must be strictly increasing indices into the "
                         "n_scan x n_scan launch lattice")
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dt))).to("cuda")  # noqa: E731
    d_x, d_y, d_a, d_p = (up(v, np.float64) for v in (sb.x_sim, sb.y_sim, sb.a_sim, sb.p_sim))
    d_idx = up(idx, np.int64)
    nbytes = int(lib.sosat_sim2d_workspace_bytes(n, int(n_scan)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    scal = torch.empty(8, dtype=torch.float64, device="cuda")
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(lib.sosat_sim2d_prepare(ptr(d_x), ptr(d_y), ptr(d_a), ptr(d_p), ptr(d_idx), n,
                                       int(n_scan), SIM2D_R_SELECT, SIM2D_SWEEPS, ptr(ws), nbytes,
                                       ptr(scal), _stream(torch)))
    h = _lib.to_numpy(scal)  # beam centre, bounding box of the selected rays, their number
    if h[6] < 3:
        raise ValueError("sim2d: fewer than 3 rays within 21 cm of the beam centre")
    grid_x, grid_y = np.mgrid[h[2]:h[3]:complex(0, SIM2D_GRID), h[4]:h[5]:complex(0, SIM2D_GRID)]
    d_gx, d_gy = up(grid_x[:, 0], np.float64), up(grid_y[0, :], np.float64)
    out_a = torch.empty((SIM2D_GRID, SIM2D_GRID), dtype=torch.float64, device="cuda")
    out_p = torch.empty_like(out_a)
    _lib.check(lib.sosat_sim2d_eval(ptr(d_x), ptr(d_y), ptr(d_a), ptr(d_p), n, int(n_scan), ptr(ws),
                                    ptr(d_gx), ptr(d_gy), SIM2D_GRID, SIM2D_R_ZERO, ptr(out_a),
                                    ptr(out_p), _stream(torch)))
    return _Grid2d(grid_x, grid_y, _lib.to_numpy(out_a), _lib.to_numpy(out_p))


class _StageData:
    def __init__(self, x_data, y_data, a_data, p_data):
        self.a_data, self.p_data, self.x_data, self.y_data = a_data, p_data, x_data, y_data


def _bspline_basis(t, xs, k=3):
    """B-spline basis functions of degree ``k`` on the knot vector ``t`` at ``xs`` (Cox-de Boor):
    ``[len(xs), len(t) - k - 1]``; the last interval is closed on the right."""
    t = np.asarray(t, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    nb = len(t) - k - 1
    span = np.clip(np.searchsorted(t, xs, side="right") - 1, k, nb - 1)  # t[span] <= x < t[span+1]
    B = np.zeros((len(xs), len(t) - 1))
    B[np.arange(len(xs)), span] = 1.0
    x = xs[:, None]
    for d in range(1, k + 1):
        m = len(t) - 1 - d                      # basis functions of degree d
        left = t[d:d + m] - t[:m]
        right = t[d + 1:d + 1 + m] - t[1:1 + m]
        with np.errstate(divide="ignore", invalid="ignore"):
            wl = np.where(left > 0, (x - t[:m]) / left, 0.0)
            wr = np.where(right > 0, (t[d + 1:d + 1 + m] - x) / right, 0.0)
        B = wl * B[:, :m] + wr * B[:, 1:m + 1]
    return B[:, :nb]


def _cubic_spline_operator(x, x_new):
    """``M`` with ``s(x_new) = M @ z`` for the interpolating cubic spline through ``(x, z)`` that
    FITPACK builds for ``interp2d(kind="cubic")`` / ``RectBivariateSpline(kx=3, s=0)``: knots
    ``[x0]*4 + x[2:-2] + [x_end]*4`` (not-a-knot), arguments outside ``[x0, x_end]`` clamped to
    the ends (fpbisp).  Index bookkeeping on ~100 numbers, done on the host like the twiddles of
    the far field; the field data are contracted with it on the device (``sosat_sim_data``)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1 or len(x) < 4 or not np.all(np.diff(x) > 0):
        raise ValueError("sim_data needs at least 4 strictly increasing grid coordinates per axis")
    t = np.concatenate([[x[0]] * 4, x[2:-2], [x[-1]] * 4])
    A = _bspline_basis(t, x)
    E = _bspline_basis(t, np.clip(np.asarray(x_new, dtype=np.float64), x[0], x[-1]))
    return np.ascontiguousarray(np.linalg.solve(A.T, E.T).T)


def sim_data(sb_in, STAGE_RANGE, STEP):
    """Resample the ``sim2d`` grid onto the holography stage grid (ray_trace.py:1720-1754) on the
    GPU.  The reference calls ``scipy.interpolate.interp2d(kind='cubic')`` (removed in SciPy
    1.14) on ``|b| / max|b|`` and ``arg b``; that is FITPACK's interpolating bicubic spline, a
    separable linear map ``My z Mx^T`` of the grid values, contracted here in fp64 on the device."""
    lib = _lib.require_gpu()
    torch = _torch()
    a = np.ascontiguousarray(np.asarray(sb_in.a_sim, dtype=np.float64))
    p = np.ascontiguousarray(np.asarray(sb_in.p_sim, dtype=np.float64))
    if a.ndim != 2 or a.shape[0] != a.shape[1] or a.shape != p.shape:
        raise ValueError("sim_data expects square [n, n] a_sim / p_sim grids of equal shape")
    n = a.shape[0]
    x_interp = np.asarray(sb_in.x_sim)[:, 0]
    y_interp = np.asarray(sb_in.y_sim)[0, :]
    x_data = np.arange(-STAGE_RANGE / 2, STAGE_RANGE / 2, STEP)
    y_data = np.arange(-STAGE_RANGE / 2, STAGE_RANGE / 2, STEP)
    m = len(x_data)
    Mx = _cubic_spline_operator(x_interp, x_data)
    My = _cubic_spline_operator(y_interp, y_data)
    up = lambda v: torch.from_numpy(v).to("cuda")  # noqa: E731
    d_a, d_p, d_Mx, d_My = up(a), up(p), up(Mx), up(My)
    a_data = torch.empty((m, m), dtype=torch.float64, device="cuda")
    p_data = torch.empty_like(a_data)
    work = torch.empty(int(lib.sosat_sim_data_work_doubles(n, m)), dtype=torch.float64, device="cuda")
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(lib.sosat_sim_data(ptr(d_a), ptr(d_p), n, ptr(d_Mx), ptr(d_My), m, ptr(a_data),
                                  ptr(p_data), ptr(work), _stream(torch)))
    x_grid, y_grid = np.meshgrid(x_data, y_data)
    return _StageData(x_grid, y_grid, _lib.to_numpy(a_data), _lib.to_numpy(p_data))


# --- fused trace + binning (north_star) -------------------------------------------------------
class BinnedField:
    """Result of :func:`near_field_binned`.

    ``b``      complex64 ``[n_rx, n_freq, N, N]`` mean of ``a exp(i p)`` per cell (0 where empty),
               phase referred to the mean optical path of the valid rays;
    ``count``  float32 ``[n_rx, N, N]`` rays per cell;
    ``stats``  float64 ``[n_rx, 8]`` (``_lib.STAT_*``);
    ``x_cm`` / ``y_cm``  cell-centre coordinates (x_sim / y_sim convention, cm).
    Arrays are numpy unless ``device=True`` was requested (then torch CUDA tensors).
    """

    def __init__(self, b, count, stats, extent_cm, grid_n, re=None, im=None):
        self.b, self.count, self.stats = b, count, stats
        self.extent_cm, self.grid_n = extent_cm, grid_n
        self.re, self.im = re, im
        edges = (np.arange(grid_n) + 0.5) * (extent_cm / grid_n) - extent_cm / 2
        self.x_cm = edges
        self.y_cm = edges.copy()

    @property
    def mean_opl(self):
        return self.stats[:, _lib.STAT_SUM_OPL] / np.maximum(self.stats[:, _lib.STAT_N_VALID], 1)


def trace_bin_device(tele_geo, rx_list, freq_specs, grid_n, extent_mm, *, opl_ref=0.0,
                     row_begin=0, row_end=None, planes=None, stats=None, bins=True, tile_rows=None):
    """Launch ``sosat_trace_bin`` on the current device/stream and return the raw accumulators
    ``(re, im, cnt, stats)`` as CUDA tensors (accumulated into ``planes`` / ``stats`` when
    given, so several row blocks -- or several GPUs after an all-reduce -- add up).
    ``tile_rows``: instead of the row block, a list (or CUDA int32 tensor) of 16-row tile rows of
    the launch grid to trace (``sosat_trace_bin_tiles``; ``dist.partition_tile_rows``)."""
    lib = _lib.require_gpu()
    torch = _torch()
    n_scan = int(tele_geo.n_scan)
    if row_end is None:
        row_end = n_scan
    rx = np.ascontiguousarray(np.asarray(rx_list, dtype=np.float64).reshape(-1, 3))
    n_rx = rx.shape[0]
    freqs = geometry.make_freqs(freq_specs)
    n_freq = len(freqs)
    _upload_geometry(lib, torch, tele_geo)
    d_rx = torch.from_numpy(rx).to("cuda")
    if stats is None:
        stats = torch.zeros((n_rx, _lib.NUM_STATS), dtype=torch.float64, device="cuda")
    if bins:
        if planes is None:
            re = torch.zeros((n_rx, n_freq, grid_n, grid_n), dtype=torch.float32, device="cuda")
            im = torch.zeros_like(re)
            cnt = torch.zeros((n_rx, grid_n, grid_n), dtype=torch.float32, device="cuda")
        else:
            re, im, cnt = planes
        ptrs = [ctypes.c_void_p(t.data_ptr()) for t in (re, im, cnt)]
    else:
        re = im = cnt = None
        ptrs = [ctypes.c_void_p(0)] * 3
    if tile_rows is not None:
        if not hasattr(tile_rows, "is_cuda"):
            tile_rows = torch.from_numpy(np.ascontiguousarray(np.asarray(tile_rows, dtype=np.int32))).to("cuda")
        if tile_rows.dtype != torch.int32 or not tile_rows.is_cuda or not tile_rows.is_contiguous():
            raise ValueError("tile_rows must be a contiguous CUDA int32 tensor (or a host sequence)")
        _lib.check(lib.sosat_trace_bin_tiles(
            ctypes.c_void_p(d_rx.data_ptr()), n_rx, freqs, n_freq, n_scan,
            ctypes.c_void_p(tile_rows.data_ptr()), int(tile_rows.numel()), int(grid_n),
            float(extent_mm), float(opl_ref), *ptrs, ctypes.c_void_p(stats.data_ptr()), _stream(torch)))
        return re, im, cnt, stats
    _lib.check(lib.sosat_trace_bin(
        ctypes.c_void_p(d_rx.data_ptr()), n_rx, freqs, n_freq, n_scan, row_begin, row_end,
        int(grid_n), float(extent_mm), float(opl_ref), *ptrs, ctypes.c_void_p(stats.data_ptr()),
        _stream(torch)))
    return re, im, cnt, stats


def bins_to_field_device(re, im, cnt, stats, freq_specs, opl_ref=0.0, *, sync: bool = True):
    """Normalise the accumulators to the mean field per cell and refer the phase to the mean
    OPL of the valid rays (the reference's ``OPL - mean(OPL)``, ray_trace.py:1601).

    ``sync=False``: the mean OPL is read from ``stats`` ON THE DEVICE
    (``sosat_bins_to_field_stats``), nothing is copied to the host and ``(b, None)`` is
    returned, so trace -> field -> far field is one uninterrupted stream of launches."""
    lib = _lib.require_gpu()
    torch = _torch()
    n_rx, n_freq, n, _ = re.shape
    b = torch.empty((n_rx, n_freq, n, n, 2), dtype=torch.float32, device="cuda")
    if not sync:
        # one launch for the whole stack of (receiver, frequency) fields
        ks = (ctypes.c_double * n_freq)(*[geometry.scalar(k) / 1e3 / 2 for k, _, _ in freq_specs])
        _lib.check(lib.sosat_bins_to_field_stats_batch(
            ctypes.c_void_p(re.data_ptr()), ctypes.c_void_p(im.data_ptr()),
            ctypes.c_void_p(cnt.data_ptr()), n * n, n_rx, n_freq, ctypes.c_void_p(stats.data_ptr()),
            ks, float(opl_ref), ctypes.c_void_p(b.data_ptr()), _stream(torch)))
        return torch.view_as_complex(b), None
    h_stats = _lib.to_numpy(stats)
    for i in range(n_rx):
        nv = h_stats[i, _lib.STAT_N_VALID]
        mean_opl = h_stats[i, _lib.STAT_SUM_OPL] / nv if nv > 0 else 0.0
        for f, (k, _, _) in enumerate(freq_specs):
            phi0 = geometry.scalar(k) / 1e3 / 2 * (mean_opl - opl_ref)
            phi0 = float(np.mod(phi0, 2 * np.pi))
            _lib.check(lib.sosat_bins_to_field(
                ctypes.c_void_p(re[i, f].data_ptr()), ctypes.c_void_p(im[i, f].data_ptr()),
                ctypes.c_void_p(cnt[i].data_ptr()), n * n, phi0,
                ctypes.c_void_p(b[i, f].data_ptr()), _stream(torch)))
    return torch.view_as_complex(b), h_stats


def near_field_binned(tele_geo, rx, grid_n: int = 256, extent_cm: Optional[float] = None,
                      freqs: Optional[Sequence] = None, *, device: bool = False,
                      keep_sums: bool = False) -> BinnedField:
    """Trace the ray bundle of one or several feeds and bin it into the complex near field
    ``b(x, y)`` on a regular ``grid_n x grid_n`` grid over ``[-extent/2, extent/2)^2`` of the
    source plane (default extent: ``tele_geo.diam`` = 42 cm, the stop diameter).

    ``rx``     one position ``[x, y, z]`` mm or a sequence of them (focal-plane sweep);
    ``freqs``  ``None`` -> the frequency currently set on ``tele_geo``; else a sequence of
               ``(k, th_fwhp_x, th_fwhp_y)`` (see geometry.freq_specs_from_table).  The ray
               geometry is frequency independent, so all frequencies share one trace.
    """
    if extent_cm is None:
        extent_cm = geometry.scalar(getattr(tele_geo, "diam", 0.42)) * 100.0
    specs = list(freqs) if freqs is not None else [
        (tele_geo.k, tele_geo.th_fwhp_x, tele_geo.th_fwhp_y)]
    out_b, out_c, out_s, out_re, out_im = [], [], [], [], []
    rx_arr = np.asarray(rx, dtype=np.float64).reshape(-1, 3)
    for f0 in range(0, len(specs), _lib.MAX_FREQS):
        chunk = specs[f0:f0 + _lib.MAX_FREQS]
        re, im, cnt, stats = trace_bin_device(tele_geo, rx_arr, chunk, grid_n, extent_cm * 10.0)
        b, _ = bins_to_field_device(re, im, cnt, stats, chunk, sync=False)  # one launch, no sync
        out_b.append(b)
        out_c, out_s = cnt, stats
        if keep_sums:
            out_re.append(re)
            out_im.append(im)
    torch = _torch()
    out_s = _lib.to_numpy(out_s)  # the only read-back of the statistics (identical for every chunk)
    b = torch.cat(out_b, dim=1) if len(out_b) > 1 else out_b[0]
    re = torch.cat(out_re, dim=1) if keep_sums else None
    im = torch.cat(out_im, dim=1) if keep_sums else None
    if not device:
        b, out_c = _lib.to_numpy(b), _lib.to_numpy(out_c)
        re = _lib.to_numpy(re) if keep_sums else None
        im = _lib.to_numpy(im) if keep_sums else None
    return BinnedField(b, out_c, out_s, extent_cm, grid_n, re, im)


SWEEP_BYTES_PER_PASS = 12e9   # device memory budget of one pass of beam_sweep (fields + workspace)


def beam_sweep(tele_geo, rx_list, freqs, grid_n: int = 256, extent_cm: Optional[float] = None,
               pad: int = 0):
    """BASELINE configs[3]: near + far field of every (receiver, frequency) pair of a focal-plane
    sweep, on the device end to end -- ONE fused trace+bin launch for all receivers and
    frequencies, ONE batched far-field transform for all fields, and the far-field FWHM search
    per field; only scalars come back.

    ``freqs``: sequence of ``(k, th_fwhp_x, th_fwhp_y)``; ``pad``: zero cells added on every side
    of the near field before the transform (finer angle sampling, like ``zero_pad``).  Returns a
    dict of arrays ``[n_rx, n_freq]``: ``fwhm_arcmin`` (``ff_fwhm_est`` on the true FFT-bin
    angles), ``peak`` (|B| at the beam peak), ``peak_index`` ``[n_rx, n_freq, 2]``, and per
    receiver ``n_valid``, ``n_binned``, ``mean_opl``."""
    from . import opt_analyze
    torch = _torch()
    if extent_cm is None:
        extent_cm = geometry.scalar(getattr(tele_geo, "diam", 0.42)) * 100.0
    specs = list(freqs)
    if len(specs) > _lib.MAX_FREQS:
        raise ValueError(f"at most {_lib.MAX_FREQS} frequencies per sweep call")
    rx_arr = np.asarray(rx_list, dtype=np.float64).reshape(-1, 3)
    n_freq, n_far = len(specs), grid_n + 2 * pad
    # receivers per pass: the far-field stack and its workspace (~80 B per padded cell and field)
    # stay below ~12 GB however long the receiver list is
    per_pass = max(1, min(int(SWEEP_BYTES_PER_PASS // (80 * n_far * n_far * n_freq)), 4096 // n_freq))
    parts = []
    for r0 in range(0, rx_arr.shape[0], per_pass):
        rxs = rx_arr[r0:r0 + per_pass]
        re, im, cnt, stats = trace_bin_device(tele_geo, rxs, specs, grid_n, extent_cm * 10.0)
        b, _ = bins_to_field_device(re, im, cnt, stats, specs, sync=False)
        del re, im, cnt
        fields = b.reshape(len(rxs) * n_freq, grid_n, grid_n)
        B = opt_analyze.far_field_batch(fields, device=True, pad=pad)  # the zeros are never stored
        del fields, b
        parts.append(_sweep_scalars(B, _lib.to_numpy(stats), specs, len(rxs),
                                    extent_cm / grid_n / 100.0))
        del B
    if len(parts) == 1:
        return parts[0]
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def _sweep_scalars(B, stats, specs, n_rx, dx_m):
    """Far-field FWHM / peak of every field of a sweep: ONE batched search, ONE read-back."""
    n_freq, n = len(specs), int(B.shape[-1])
    idx, val = ff_fwhm_indices_device(B)
    idx = idx.reshape(n_rx, n_freq, 6).astype(np.int64)
    lam = np.array([2 * np.pi / geometry.scalar(k) for k, _, _ in specs])
    dth = lam / (n * dx_m) * (180 * 60 / np.pi)  # arcmin per far-field cell
    fwhm = 0.5 * (np.abs(idx[..., 3] - idx[..., 2]) + np.abs(idx[..., 5] - idx[..., 4])) * dth[None, :]
    st = np.asarray(stats)
    return {"fwhm_arcmin": fwhm, "peak": np.sqrt(val.astype(np.float64)).reshape(n_rx, n_freq),
            "peak_index": idx[..., :2].copy(),
            "n_valid": st[:, _lib.STAT_N_VALID].copy(), "n_binned": st[:, _lib.STAT_N_BINNED].copy(),
            "mean_opl": st[:, _lib.STAT_SUM_OPL] / np.maximum(st[:, _lib.STAT_N_VALID], 1)}


def trace_stats(tele_geo, rx):
    """Statistics-only pass (no binning): per-feed ``[n_valid, sum OPL, sum exit dir, ...]``."""
    rx_arr = np.asarray(rx, dtype=np.float64).reshape(-1, 3)
    spec = [(tele_geo.k, tele_geo.th_fwhp_x, tele_geo.th_fwhp_y)]
    _, _, _, stats = trace_bin_device(tele_geo, rx_arr, spec, 0, 0.0, bins=False)
    return _lib.to_numpy(stats)
