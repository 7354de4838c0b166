"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the sosat_optics beam-simulation hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``sosat_optics_b200``) never does, and fails loudly without its CUDA library.

What is restated here (reference file:line, relative to the reference checkout):

* ray tracer ``rx_to_lyot``            ray_trace.py:1070-1591  -> C, ``trace_oracle.c``
* ``getNearField``                     ray_trace.py:1594-1614  -> :func:`get_near_field`
* ``get_beam_cent/get_fwhm/get_angle_out`` ray_trace.py:1652-1672
* ``ff_fwhm_est``                      ray_trace.py:1757-1779
* ``zero_pad``                         opt_analyze.py:20-32
* ``a2b``                              opt_analyze.py:220-230  (numpy.fft, pocketfft)
* ``coords_spat_to_ang``               opt_analyze.py:274-299
* ``b2a``                              opt_analyze.py:208-217
* ``beam_convolve``                    opt_analyze.py:60-88
* ``coords_ang_to_spat``               opt_analyze.py:233-271
* ``ghz_to_m`` / ``m_to_ghz``          opt_analyze.py:42-57
* near-field binning b(x,y): no reference function exists;
  :func:`bin_near_field` is the specification the CUDA kernel is held to.

Parity pin: the reference ships no tests.  This oracle is pinned against golden vectors
produced by running the reference itself in the dev container (``oracle/make_golden.py``
-> ``tests/golden/*.npz``) and against the numbers stored in the reference notebooks
(FWHM 11.06 / 13.32 cm, 21.50 arcmin, ...).  See tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "libsosat_oracle.so")

# --- constants restated from ot_geo.py -------------------------------------------------
N_SI = 3.416  # ot_geo.py:13
N_VAC = 1.0  # ot_geo.py:14
LENS_T1, LENS_T2, LENS_T3 = 4.5633e1, 4.9101e1, 3.1790e1  # ot_geo.py:39-41
LENS3_Y = (1.5 + 0.45 + 2.1553254002897502) * 1e1  # ot_geo.py:43
LENS2_Y = LENS3_Y + 10.676699903527055e1 + LENS_T3  # ot_geo.py:44
LENS1_Y = LENS2_Y + 3.8933e1 + +0.45e1 + 0.5e1 + 0.45e1 + 47.776194563951336e1 + LENS_T2
Y_LYOT = LENS1_Y + 1.0553506281444689e1 + LENS_T1  # ot_geo.py:46
TH = np.pi / 2  # ot_geo.py:32-37

# trace order 3b, 3a, 2b, 2a, 1b, 1a: (name, c, k, a1, a2, a3, y0, n1, n2, t_lo, t_hi)
SURFACES = (
    ("3b", -2.50138371e-02, -18.18966463, 2.29039839e-03, 5.00824363e-06, 1.94511404e-09,
     LENS3_Y, N_VAC, N_SI, 2.0, 1600.0),  # ot_geo.py:342-346, ray_trace.py:1130
    ("3a", -2.48540190e-02, -30.00640716, 4.05886709e-04, 5.00081009e-06, 5.00669812e-09,
     LENS3_Y + 3.1790e1, N_SI, N_VAC, 5.0, 200.0),  # ot_geo.py:320-324,441; :1198
    ("2b", 2.39616010e-02, -1.04603079, -8.38235791e-03, -1.39048036e-06, 1.84669164e-10,
     LENS2_Y, N_VAC, N_SI, 10.0, 600.0),  # ot_geo.py:215-219; :1261
    ("2a", 1.22111813e-02, -30.00046667, 1.56460553e-03, 3.06349779e-06, -5.00246955e-09,
     LENS2_Y + 4.9101e1, N_SI, N_VAC, 5.0, 400.0),  # ot_geo.py:191-195,305; :1326
    ("1b", -5.97879325e-3, 29.99835006, 6.86096781e-3, 4.44719534e-7, -2.62178001e-9,
     LENS1_Y, N_VAC, N_SI, 5.0, 1000.0),  # ot_geo.py:57-61; :1393
    ("1a", 9.65674282e-3, -30.00122787, 1.45021174e-3, 2.84858123e-6, -5.02176737e-9,
     LENS1_Y + LENS_T1, N_SI, N_VAC, 5.0, 220.0),  # ot_geo.py:79-83,174; :1460
)


class _Surface(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("c", "k", "a1", "a2", "a3", "y0", "th", "n1", "n2", "t_lo", "t_hi")]


class _Geom(ctypes.Structure):
    _fields_ = [("surf", _Surface * 6)] + [(n, ctypes.c_double) for n in
                ("y_lyot", "y_source", "fwhp_x", "fwhp_y", "clip_l3", "clip_stop", "cone")]


def build(force: bool = False) -> str:
    """Compile trace_oracle.c -> oracle/_build/libsosat_oracle.so (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "trace_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(_BUILD, exist_ok=True)
    # -ffp-contract=off: keep the reference's operation order (no fused multiply-add).
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o",
           _LIB_PATH, src, "-lm"]
    subprocess.check_call(cmd)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        L = ctypes.CDLL(path)
        L.oracle_rx_to_lyot.restype = ctypes.c_long
        L.oracle_rx_to_lyot.argtypes = [
            ctypes.POINTER(_Geom), ctypes.POINTER(ctypes.c_double), ctypes.c_long,
            ctypes.c_long, ctypes.c_long, ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_long), ctypes.c_int]
        L.oracle_brentq_surface.restype = ctypes.c_double
        L.oracle_brentq_surface.argtypes = [
            ctypes.POINTER(_Surface), ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double,
            ctypes.POINTER(ctypes.c_int)]
        L.oracle_root_fn.restype = ctypes.c_double
        L.oracle_root_fn.argtypes = [
            ctypes.POINTER(_Surface), ctypes.POINTER(ctypes.c_double),
            ctypes.POINTER(ctypes.c_double), ctypes.c_double]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _scalar(v) -> float:
    """SatGeo attributes may be python floats or shape-(1,1) arrays."""
    return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


def default_tele_geo(**over):
    """A SatGeo look-alike with the class defaults of ot_geo.py:8-29."""
    g = SimpleNamespace(
        n_si=N_SI, n_vac=N_VAC, th_fwhp_x=np.deg2rad(35), th_fwhp_y=np.deg2rad(35),
        n_scan=100, lambda_=(30.0 / 150.0) * 0.01, diam=0.42, x_ap=0, y_ap=0, z_ap=0,
        y_source=0)
    g.k = 2 * np.pi / g.lambda_
    for k, v in over.items():
        setattr(g, k, v)
    return g


def _surface_struct(i: int, n_si: float, n_vac: float) -> _Surface:
    _, c, k, a1, a2, a3, y0, n1, n2, lo, hi = SURFACES[i]
    n1 = n_si if n1 == N_SI else n_vac
    n2 = n_si if n2 == N_SI else n_vac
    return _Surface(c, k, a1, a2, a3, y0, TH, n1, n2, lo, hi)


def make_geom(tele_geo) -> _Geom:
    g = _Geom()
    for i in range(6):
        g.surf[i] = _surface_struct(i, _scalar(tele_geo.n_si), _scalar(tele_geo.n_vac))
    g.y_lyot = Y_LYOT
    g.y_source = _scalar(tele_geo.y_source)
    g.fwhp_x = _scalar(tele_geo.th_fwhp_x)
    g.fwhp_y = _scalar(tele_geo.th_fwhp_y)
    g.clip_l3 = 392 / 2
    g.clip_stop = 210.0
    g.cone = 0.4
    return g


def rx_to_lyot(P_rx, tele_geo, plot=False, col="b", *, ray_begin=0, ray_end=None,
               n_threads=0, return_status=False):
    """Restatement of ray_trace.py:1070-1591.  Returns out[17, n] float64."""
    n_scan = int(tele_geo.n_scan)
    n_all = n_scan * n_scan
    if ray_end is None:
        ray_end = n_all
    n = ray_end - ray_begin
    out = np.zeros((17, n), dtype=np.float64)
    status = np.zeros(n, dtype=np.int32)
    stats = (ctypes.c_long * 1)(0)
    rx = np.asarray(P_rx, dtype=np.float64).reshape(3).copy()
    g = make_geom(tele_geo)
    nerr = lib().oracle_rx_to_lyot(
        ctypes.byref(g), rx.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_scan,
        ray_begin, ray_end, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), stats, int(n_threads))
    if nerr:
        # the reference's brentq raises "f(a) and f(b) must have different signs"
        raise ValueError(f"{nerr} rays: f(a) and f(b) must have different signs")
    if return_status:
        return out, status, int(stats[0])
    return out


def get_near_field(tele_geo, rx, out=None):
    """ray_trace.py:1594-1614.  ``out`` may be supplied (e.g. the GPU result) to test only
    the post-processing."""
    if out is None:
        out = rx_to_lyot(rx, tele_geo)
    xx = np.where(out[4] != 0)
    a_sim = out[4][xx]
    k = _scalar(tele_geo.k)
    p_sim = np.mod(k * (out[3][xx] - np.mean(out[3][xx])) / 1e3 / 2, 2 * np.pi)
    x_sim = out[0][xx] / 1e1
    y_sim = out[2][xx] / 1e1
    tan_og = np.array([np.mean(out[8][xx]), np.mean(out[9][xx]), np.mean(out[10][xx])])
    return SimpleNamespace(x_sim=x_sim, y_sim=y_sim, a_sim=a_sim,
                           p_sim=p_sim - np.mean(p_sim), tan_og=tan_og)


def get_beam_cent(sb):  # ray_trace.py:1652-1654
    return [np.mean(sb.x_sim), np.mean(sb.y_sim)]


def get_fwhm(sb):  # ray_trace.py:1657-1666
    bc = get_beam_cent(sb)
    sel = np.where(sb.a_sim ** 2 > np.max(sb.a_sim ** 2) / 2)
    return abs(sb.y_sim - bc[1])[sel].max(), abs(sb.x_sim - bc[0])[sel].max()


def get_angle_out(sb):  # ray_trace.py:1669-1672
    theta_x = abs(90 - np.rad2deg(np.arctan2(sb.tan_og[1], sb.tan_og[2])))
    theta_y = abs(90 - np.rad2deg(np.arctan2(sb.tan_og[1], sb.tan_og[0])))
    return [theta_x, theta_y]


def ff_fwhm_est(beam_fft, x_ang, y_ang):  # ray_trace.py:1757-1779
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


def ghz_to_m(freq_ghz):  # opt_analyze.py:42-48
    return (3 * 10 ** 8) / (freq_ghz * 1e9)


def m_to_ghz(wavelength):  # opt_analyze.py:51-57
    return ((3 * 10 ** 8) / wavelength) / 1e9


def zero_pad(x_in, y_in, beam_in, pts):  # opt_analyze.py:20-32
    x_int = abs(x_in[0, 0] - x_in[1, 1])
    y_int = abs(y_in[0, 0] - y_in[1, 1])
    beam_out = np.pad(beam_in, pts, mode="constant")
    x_new = np.array(np.arange(len(beam_out))) * x_int
    y_new = np.array(np.arange(len(beam_out))) * y_int
    x_new -= np.mean(x_new)
    y_new -= np.mean(y_new)
    x_new, y_new = np.meshgrid(x_new, y_new)
    beam_out = np.where(np.isnan(beam_out), 0, beam_out)
    return x_new + np.mean(x_in), y_new + np.mean(y_in), beam_out.T


def a2b(apert, phi):  # opt_analyze.py:220-230
    apert = abs(apert) * np.exp(phi * np.pi / 180.0 * complex(0, 1))
    beam_complex = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(apert)))
    phase = np.arctan2(np.imag(beam_complex), np.real(beam_complex))
    return beam_complex, phase


def coords_spat_to_ang(x, y, freq):  # opt_analyze.py:274-299
    lam = (3 * 10 ** 8) / (freq * 1e9)
    delta_x = abs(np.max(x) - np.min(x)) / (len(x) - 1)
    delta_y = abs(np.max(y) - np.min(y)) / (len(y) - 1)
    delta_th = (lam / delta_x) / len(x)
    delta_ph = (lam / delta_y) / len(y)
    x_ang = np.linspace(-int(len(x) / 2), int(len(x) / 2), int(len(x))) * delta_th
    y_ang = np.linspace(-int(len(x) / 2), int(len(x) / 2), int(len(x))) * delta_ph
    return tuple(np.meshgrid(x_ang, y_ang))


def b2a(beam, phase):  # opt_analyze.py:208-217
    beam_complex = abs(beam) * np.exp(phase * np.pi / 180.0 * complex(0, 1))
    # line 213 computes fftshift(beam_complex) and line 214 discards it: the input is NOT shifted
    aper_field = np.fft.fftshift(np.fft.ifft2(beam_complex))
    return aper_field, np.arctan2(np.imag(aper_field), np.real(aper_field))


def horn_aperture(x, y, apert1, apert2):  # opt_analyze.py:65-73
    if apert1 < apert2:
        disc = np.cos(np.pi * y / apert2)
        disc = np.where(abs(y) <= apert2 / 2, disc, 0)
        disc = np.where(abs(x) <= apert1 / 2, disc, 0)
    else:
        disc = np.cos(np.pi * x / apert1)
        disc = np.where(abs(x) <= apert1 / 2, disc, 0)
        disc = np.where(abs(y) <= apert2 / 2, disc, 0)
    return disc


def beam_convolve(x, y, beam, apert1, apert2, plots=0):  # opt_analyze.py:60-88
    disc = horn_aperture(x, y, apert1, apert2)
    disc_fft = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(disc)))
    beam_fft = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(beam)))
    beam_conv = beam_fft * disc_fft
    beam_final = np.fft.ifftshift(np.fft.ifft2(np.fft.ifftshift(beam_conv)))
    return x, y, beam_final


def coords_ang_to_spat(theta_x, theta_y, freq):  # opt_analyze.py:233-271
    lam = (3 * 10 ** 8) / (freq * 1e9)
    delta_th = abs(np.max(theta_y) - np.min(theta_y)) / (len(theta_y) - 1)  # :241 overwrites :238
    delta_x = (lam / delta_th) / len(theta_x)
    delta_y = (lam / delta_th) / len(theta_y)
    half, cnt = int(len(theta_x) / 2), int(len(theta_x))
    x_spat = np.linspace(-half, half, cnt) * delta_x * 1e2
    y_spat = np.linspace(-half, half, cnt) * delta_y * 1e2
    return tuple(np.meshgrid(x_spat, y_spat))


def dft2_kind(b, kind):
    """The three centred transforms of opt_analyze.py as dense contractions W b W^T with
    W[k, m] = exp(sign 2 pi i (k - so)(m - si) / N) (exact integer phase) and a scale:
    0 forward  fftshift(fft2(fftshift(b)))     so = floor(N/2), si = ceil(N/2), sign -, 1
    1 inverse  ifftshift(ifft2(ifftshift(b)))  so = ceil(N/2),  si = floor(N/2), sign +, 1/N^2
    2 b2a      fftshift(ifft2(b))              so = floor(N/2), si = 0,          sign +, 1/N^2"""
    n = b.shape[0]
    so, si, sign, scale = {0: (n // 2, (n + 1) // 2, -1, 1.0),
                           1: ((n + 1) // 2, n // 2, 1, 1.0 / n ** 2),
                           2: (n // 2, 0, 1, 1.0 / n ** 2)}[kind]
    k = np.arange(n)[:, None] - so
    m = np.arange(n)[None, :] - si
    w = np.exp(sign * 2j * np.pi * ((k * m) % n) / n)
    return (w @ b @ w.T) * scale


def dft2_centered(b):
    """The a2b transform written as the dense contraction W_r b W_c^T:
    W[k, n] = exp(-2 pi i (k - floor(N/2)) (n - ceil(N/2)) / N), exact integer phase."""
    def w(n):
        k = np.arange(n)[:, None] - n // 2
        m = np.arange(n)[None, :] - (n + 1) // 2
        ph = (k * m) % n
        return np.exp(-2j * np.pi * ph / n)
    return w(b.shape[0]) @ b @ w(b.shape[1]).T


# --- near-field binning specification -------------------------
def ray_phase(out, k, opl_ref):
    """Per-ray phase before any wrap: k (OPL - opl_ref) / 1e3 / 2 (ray_trace.py:1601)."""
    return k * (out[3] - opl_ref) / 1e3 / 2


def bin_near_field(out, k, grid_n, extent_mm, opl_ref):
    """Sum a*exp(i p) and the ray count per cell of a grid_n x grid_n grid covering
    [-extent/2, extent/2)^2 mm in (x, z) of the source plane; row index = z (the
    reference's y_sim), column index = x (x_sim).  Valid rays only (a != 0)."""
    sel = out[4] != 0
    x, z, a = out[0][sel], out[2][sel], out[4][sel]
    p = k * (out[3][sel] - opl_ref) / 1e3 / 2
    inv = grid_n / extent_mm
    ix = np.floor((x + extent_mm / 2) * inv).astype(np.int64)
    iz = np.floor((z + extent_mm / 2) * inv).astype(np.int64)
    ok = (ix >= 0) & (ix < grid_n) & (iz >= 0) & (iz < grid_n)
    re = np.zeros((grid_n, grid_n))
    im = np.zeros((grid_n, grid_n))
    cnt = np.zeros((grid_n, grid_n))
    np.add.at(re, (iz[ok], ix[ok]), a[ok] * np.cos(p[ok]))
    np.add.at(im, (iz[ok], ix[ok]), a[ok] * np.sin(p[ok]))
    np.add.at(cnt, (iz[ok], ix[ok]), 1.0)
    sums = dict(n_valid=int(sel.sum()), sum_opl=float(out[3][sel].sum()),
                n_binned=int(ok.sum()))
    return re, im, cnt, sums


def nudft_direct(x, y, b, thx, thy, lam):
    """Direct-sum far field on irregular samples: B_j = sum_p b_p exp(-2 pi i (x_p thx_j +
    y_p thy_j)/lam); sign follows the code path (numpy FFT), not README.md:24."""
    ph = (np.outer(thx, x) + np.outer(thy, y)) / lam
    return np.exp(-2j * np.pi * ph) @ b


def sim_data(sb2_a, sb2_p, sb2_x, sb2_y, stage_range, step):
    """Glue step ray_trace.py:1720-1754 (scattered-grid -> stage grid), used only to rebuild
    the 800x800 input of the 1000^2 far-field golden from the committed 100x100 sim2d grid.
    ``interp2d(kind='cubic')`` no longer exists in SciPy; RectBivariateSpline(kx=ky=3, s=0)
    is the adapter oracle/ref_shim.py used when the golden was generated."""
    from scipy.interpolate import RectBivariateSpline
    beam_final = sb2_a.T * np.exp(complex(0, 1) * sb2_p.T)
    x_interp = sb2_x[:, 0]
    y_interp = sb2_y[0, :]

    def interp2d(z):
        spl = RectBivariateSpline(x_interp, y_interp, np.asarray(z).T, kx=3, ky=3, s=0)
        return lambda xn, yn: spl(xn, yn).T

    func_beam = interp2d(np.abs(beam_final) / np.max(np.abs(beam_final)))
    func_phase = interp2d(np.arctan2(np.imag(beam_final), np.real(beam_final)))
    x_data = np.arange(-stage_range / 2, stage_range / 2, step)
    y_data = np.arange(-stage_range / 2, stage_range / 2, step)
    beam_data = func_beam(x_data, y_data)
    phase_data = func_phase(x_data, y_data)
    x_data, y_data = np.meshgrid(x_data, y_data)
    return SimpleNamespace(a_data=beam_data / np.max(beam_data), p_data=phase_data,
                           x_data=x_data, y_data=y_data)
