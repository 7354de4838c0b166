"""Host-side marshalling: SatGeo-like objects -> the C structs of include/sosat_b200.h.

The reference passes a mutable ``SatGeo`` instance around (ot_geo.py:8-29) and reads module
constants through ``from sosat_optics.ot_geo import *`` (ray_trace.py:9).  Here both are
gathered into one POD ``sosat_geom_t`` (frequency-independent optics, uploaded to
``__constant__`` memory) and one ``sosat_freq_t`` per frequency.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

from . import _lib, ot_geo


def scalar(v) -> float:
    """SatGeo attributes arrive as python floats, numpy scalars or shape-(1,1) arrays
    (``sat_nearfield.ipynb`` cell 1 assigns ``th_fwhp_x = np.deg2rad(fwhms[np.where(...), 1])``);
    the reference relies on size-1 -> scalar conversion at ray_trace.py:1573."""
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size != 1:
        raise ValueError(f"expected a scalar or size-1 array, got shape {np.shape(v)}")
    return float(a[0])


def make_geom(tele_geo) -> _lib.Geom:
    n = {"si": scalar(getattr(tele_geo, "n_si", ot_geo.SatGeo.n_si)),
         "vac": scalar(getattr(tele_geo, "n_vac", ot_geo.SatGeo.n_vac))}
    g = _lib.Geom()
    for i, s in enumerate(ot_geo.SURFACES):
        g.surf[i] = _lib.Surface(s.c, s.k, s.a1, s.a2, s.a3, s.y0, n[s.side_in], n[s.side_out],
                                 s.t_lo, s.t_hi)
    g.y_lyot = ot_geo.y_lyot
    g.y_source = scalar(tele_geo.y_source)
    g.clip_l3 = ot_geo.CLIP_LENS3_MM
    g.clip_stop = ot_geo.CLIP_STOP_MM
    g.cone = ot_geo.CONE_HALF_ANGLE
    return g


def make_freq(tele_geo) -> _lib.Freq:
    """Feed taper widths and phase scale of the geometry's current frequency."""
    return _lib.Freq(scalar(tele_geo.th_fwhp_x), scalar(tele_geo.th_fwhp_y),
                     scalar(tele_geo.k) / 1e3 / 2)


def make_freqs(specs: Iterable) -> "ctypes.Array":
    """specs: iterable of (k [rad/m], th_fwhp_x [rad], th_fwhp_y [rad])."""
    specs = list(specs)
    if not 1 <= len(specs) <= _lib.MAX_FREQS:
        raise ValueError(f"need 1..{_lib.MAX_FREQS} frequencies per call, got {len(specs)}")
    arr = (_lib.Freq * len(specs))()
    for i, (k, fx, fy) in enumerate(specs):
        arr[i] = _lib.Freq(scalar(fx), scalar(fy), scalar(k) / 1e3 / 2)
    return arr


def ghz_to_m(freq_ghz):
    """opt_analyze.py:42-48 (c = 3e8 exactly)."""
    return (3 * 10 ** 8) / (freq_ghz * 1e9)


def freq_specs_from_table(freqs_ghz: Sequence[float], table: np.ndarray, fwhm_scale: float = 1.0):
    """The notebooks' per-frequency setup (``sat_reqs.ipynb`` cell 1): wavelength from
    ``ghz_to_m``, ``k = 2 pi / lambda``, feed FWHM from the row of ``fwhm_feedhorns.npy`` whose
    first column equals ``int(freq)`` (integer GHz 70..174 only)."""
    out = []
    for f in freqs_ghz:
        row = table[table[:, 0] == int(f)]
        if row.shape[0] != 1:
            raise KeyError(f"no feed-horn table row for int({f}) GHz")
        lam = ghz_to_m(f)
        out.append((2 * np.pi / lam, np.deg2rad(row[0, 1]) * fwhm_scale,
                    np.deg2rad(row[0, 2]) * fwhm_scale))
    return out
