"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference behind four shims.

This module exists to generate golden vectors (``oracle/make_golden.py``) and to
validate the CPU restatements in ``oracle/`` in the development container, where
``/root/reference`` is mounted.  It cannot travel to the GPU box; nothing in the
product package, ``bench.py`` or the ``-m gpu`` tests imports it.

The reference (2022, numpy<1.24 / scipy<1.14 era) no longer imports on the current
stack.  The shims change no reference arithmetic:

1. ``matplotlib`` is imported at module top of all three reference modules
   (``ot_geo.py:1-2``, ``ray_trace.py:1-2``, ``opt_analyze.py:1-2``) but is not
   installed -> stub modules.
2. SciPy >= 1.12 ``brentq`` raises ``ValueError`` when f(x) is NaN
   (``_zeros_py._wrap_nan_raise``); the SciPy the notebooks ran on did not.  A NaN
   can only reach brentq after total internal reflection has poisoned the ray
   direction (``ray_trace.py:31-33``); such rays are zeroed by the NaN-false
   comparisons at ``ray_trace.py:1569-1580`` whatever brentq returns, so the wrapper
   returns NaN.
3. ``np.complex`` was removed in numpy 1.24 (``opt_analyze.py:212, 224``).
4. ``scipy.interpolate.interp2d`` was removed in SciPy 1.14 (``ray_trace.py:1729``)
   -> RectBivariateSpline adapter (glue step only, not on the hot path).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SOSAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sosat_optics"))


class _Anything:
    """Absorbs any attribute access / call made on the matplotlib stubs."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return _Anything()


def _install_matplotlib_stub():
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
        import matplotlib.image  # noqa: F401
        return
    except Exception:
        pass
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image",
                 "matplotlib.font_manager"):
        mod = types.ModuleType(name)
        mod.__getattr__ = lambda attr, _n=name: _Anything()  # type: ignore[attr-defined]
        sys.modules[name] = mod
    sys.modules["matplotlib"].rcParams = {}  # type: ignore[attr-defined]
    sys.modules["matplotlib"].rcParams = _RcParams()  # type: ignore[attr-defined]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]  # type: ignore
    sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]  # type: ignore


class _RcParams(dict):
    def update(self, *a, **k):  # matplotlib.rcParams.update({...}) at ray_trace.py:12
        pass


_loaded = None


def load_reference():
    """Return (ot_geo, ray_trace, opt_analyze) reference modules, shimmed."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_matplotlib_stub()
    if not hasattr(np, "complex"):
        np.complex = complex  # type: ignore[attr-defined]  # shim 3
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from sosat_optics import ot_geo, opt_analyze, ray_trace  # type: ignore

    from scipy import interpolate, optimize

    real_brentq = optimize.brentq

    class _OptimizeProxy:
        """shim 2: ray_trace.optimize.brentq -> NaN instead of 'value is NaN' error."""

        def __getattr__(self, name):
            return getattr(optimize, name)

        @staticmethod
        def brentq(f, a, b, *args, **kw):
            try:
                return real_brentq(f, a, b, *args, **kw)
            except ValueError as exc:
                if "NaN" in str(exc) or "nan" in str(exc):
                    return float("nan")
                raise

    ray_trace.optimize = _OptimizeProxy()

    if not hasattr(interpolate, "interp2d") or True:
        from scipy.interpolate import RectBivariateSpline

        class _InterpolateProxy:
            """shim 4: interp2d(x, y, z[len(y), len(x)], kind='cubic') adapter."""

            def __getattr__(self, name):
                return getattr(interpolate, name)

            @staticmethod
            def interp2d(x, y, z, kind="cubic"):
                assert kind == "cubic"
                spl = RectBivariateSpline(np.asarray(x), np.asarray(y),
                                          np.asarray(z).T, kx=3, ky=3, s=0)

                def call(xn, yn):
                    # interp2d returns shape (len(yn), len(xn))
                    return spl(np.asarray(xn), np.asarray(yn)).T

                return call

        ray_trace.interpolate = _InterpolateProxy()

    # tqdm progress bars write to stderr per ray; silence them.
    class _NoBar:
        def __init__(self, *a, **k):
            pass

        def update(self, n=1):
            pass

        def close(self):
            pass

    ray_trace.tqdm = _NoBar

    _loaded = (ot_geo, ray_trace, opt_analyze)
    return _loaded


def load_feedhorn_table() -> np.ndarray:
    """``feedhorn/fwhm_feedhorns.npy``: float64 (105, 3) = [GHz, FWHM_x deg, FWHM_y deg]."""
    return np.load(os.path.join(REFERENCE_ROOT, "feedhorn", "fwhm_feedhorns.npy"))
