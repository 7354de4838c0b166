"""Telescope geometry of the SO Small Aperture Telescope -- the input structure of the hot path.

Mirrors the part of the reference's ``sosat_optics/ot_geo.py`` that the ray tracer reads:
the mutable ``SatGeo`` parameter bag (ot_geo.py:8-29), the lens placement constants
(ot_geo.py:32-46) and the asphere coefficients, which the reference keeps as literals inside
each ``z??`` / ``d_z??`` function (e.g. ot_geo.py:57-61) and which are collected here into
one table, :data:`SURFACES`, in trace order.  The ray tracer evaluates the surfaces inside the
CUDA kernel (csrc/trace_device.cuh); the reference's Python-level surface functions -- sag
``z1a .. z3b`` (ot_geo.py:49-91, 182-227, 313-354), slope ``d_z1a .. d_z3b`` (:94-143, 230-276,
364-414) and the frame maps ``tele_into_m??`` / ``m??_into_tele`` (:146-179, 279-310, 417-458)
-- are generated below from the same table so that ``from ot_geo import *`` keeps providing
them (numpy, same arithmetic order; they are not on the hot path).  Plotting helpers
(``plot_lenses``, ``filt``) are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class SatGeo:
    """SAT geometry / source parameters, same attribute names and defaults as ot_geo.py:8-29.
    Notebooks mutate an instance in place; any object exposing these attributes works."""

    n_si = 3.416
    n_vac = 1.0

    th_fwhp_x = np.deg2rad(35)
    th_fwhp_y = np.deg2rad(35)

    n_scan = 100

    lambda_ = (30.0 / 150.0) * 0.01  # [m]
    k = 2 * np.pi / lambda_

    diam = 0.42  # diameter of window [m]
    x_ap = 0
    y_ap = 0
    z_ap = 0

    y_source = 0


# tilt angles of the lens frames (ot_geo.py:32-37); the kernel relies on them being pi/2
th1_l1 = th2_l1 = th1_l2 = th2_l2 = th1_l3 = th2_l3 = np.pi / 2

# lens centre thicknesses [mm] (ot_geo.py:39-41)
lens_t1 = 4.5633e1
lens_t2 = 4.9101e1
lens_t3 = 3.1790e1

# vertex positions along the optical (tele y) axis [mm] (ot_geo.py:43-46)
lens3_y = (1.5 + 0.45 + 2.1553254002897502) * 1e1
lens2_y = lens3_y + 10.676699903527055e1 + lens_t3
lens1_y = lens2_y + 3.8933e1 + +0.45e1 + 0.5e1 + 0.45e1 + 47.776194563951336e1 + lens_t2
y_lyot = lens1_y + 1.0553506281444689e1 + lens_t1

# hard-coded in rx_to_lyot
CONE_HALF_ANGLE = 0.4  # rad, ray_trace.py:1086-1087
CLIP_LENS3_MM = 392 / 2  # ray_trace.py:1138
CLIP_STOP_MM = 210.0  # ray_trace.py:1569-1570


@dataclass(frozen=True)
class Surface:
    """One refracting surface: sag = 10 [c r^2/(1+sqrt(1-(1+k) c^2 r^2)) + a1 r^2 + a2 r^4 +
    a3 r^6] mm with r in cm; ``side_in`` / ``side_out`` name the media ('vac' or 'si')."""

    name: str
    c: float
    k: float
    a1: float
    a2: float
    a3: float
    y0: float
    side_in: str
    side_out: str
    t_lo: float
    t_hi: float


# trace order: the ray leaves the focal plane and meets lens 3 first (ray_trace.py:1114-1506)
SURFACES = (
    Surface("3b", -2.50138371e-02, -18.18966463, 2.29039839e-03, 5.00824363e-06,
            1.94511404e-09, lens3_y, "vac", "si", 2.0, 1600.0),  # ot_geo.py:342-346
    Surface("3a", -2.48540190e-02, -30.00640716, 4.05886709e-04, 5.00081009e-06,
            5.00669812e-09, lens3_y + 3.1790e1, "si", "vac", 5.0, 200.0),  # :320-324, :441
    Surface("2b", 2.39616010e-02, -1.04603079, -8.38235791e-03, -1.39048036e-06,
            1.84669164e-10, lens2_y, "vac", "si", 10.0, 600.0),  # :215-219
    Surface("2a", 1.22111813e-02, -30.00046667, 1.56460553e-03, 3.06349779e-06,
            -5.00246955e-09, lens2_y + 4.9101e1, "si", "vac", 5.0, 400.0),  # :191-195, :305
    Surface("1b", -5.97879325e-3, 29.99835006, 6.86096781e-3, 4.44719534e-7,
            -2.62178001e-9, lens1_y, "vac", "si", 5.0, 1000.0),  # :57-61
    Surface("1a", 9.65674282e-3, -30.00122787, 1.45021174e-3, 2.84858123e-6,
            -5.02176737e-9, lens1_y + lens_t1, "si", "vac", 5.0, 220.0),  # :79-83, :174
)


# ---- the reference's Python-level surface functions, generated from SURFACES --------------------
_TILT = {"1a": "th1_l1", "1b": "th1_l1", "2a": "th2_l2", "2b": "th1_l2", "3a": "th1_l3", "3b": "th2_l3"}


def _make_surface_functions(surf: Surface):
    c, k, a_1, a_2, a_3, y0 = surf.c, surf.k, surf.a1, surf.a2, surf.a3, surf.y0
    th = globals()[_TILT[surf.name]]

    def sag(x, y):
        """Surface sag [mm] at local (x, y) [mm]: 10 [c r^2 / (1 + sqrt(1 - (1+k) c^2 r^2)) +
        a1 r^2 + a2 r^4 + a3 r^6], r in cm (NaN outside the conic domain, as in the reference)."""
        x_temp = x / 1e1
        y_temp = y / 1e1
        r = np.sqrt(x_temp ** 2 + y_temp ** 2)
        amp = (c * r ** 2) / (1 + np.sqrt(1 - ((1 + k) * c ** 2 * r ** 2)))
        amp += a_1 * r ** 2 + a_2 * r ** 4 + a_3 * r ** 6
        return amp * 10

    def slope(x, y):
        """(d sag / dx, d sag / dy) at local (x, y) [mm], dimensionless."""
        x_temp = x / 1e1
        y_temp = y / 1e1
        r = np.sqrt(x_temp ** 2 + y_temp ** 2)
        root = np.sqrt(1 - ((1 + k) * c ** 2 * r ** 2))
        coeff_1 = (a_1 * 2) + (a_2 * 4 * r ** 2) + (a_3 * 6 * r ** 4)
        coeff_2 = (c * 2) / (1 + root)
        coeff_3 = (c ** 3 * (k + 1) * r ** 2) / (root * (1 + root) ** 2)
        return x_temp * (coeff_1 + coeff_2 + coeff_3), y_temp * (coeff_1 + coeff_2 + coeff_3)

    def into_tele(x, y, z):
        """Surface frame -> telescope frame: rotate by the tilt about x, shift to the vertex."""
        return x, (y * np.cos(th) - z * np.sin(th)) + y0, y * np.sin(th) + z * np.cos(th)

    def tele_into(x, y, z):
        """Telescope frame -> surface frame.  (The reference shifts ``y`` in place, which mutates
        an array argument; this version leaves its arguments alone.)"""
        y = y - y0
        return x, y * np.cos(-th) - z * np.sin(-th), y * np.sin(-th) + z * np.cos(-th)

    return sag, slope, into_tele, tele_into


for _s in SURFACES:
    _sag, _slope, _into, _from = _make_surface_functions(_s)
    for _name, _fn in ((f"z{_s.name}", _sag), (f"d_z{_s.name}", _slope),
                       (f"m{_s.name}_into_tele", _into), (f"tele_into_m{_s.name}", _from)):
        _fn.__name__ = _fn.__qualname__ = _name
        globals()[_name] = _fn
del _s, _sag, _slope, _into, _from, _name, _fn
