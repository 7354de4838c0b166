"""sosat_optics_b200 -- B200 (sm_100a) implementation of the sosat_optics beam-simulation hot
path: ray trace feed -> source plane, near-field binning, far-field transform.

Use it like the reference package::

    from sosat_optics_b200 import ot_geo, ray_trace, opt_analyze

The arithmetic lives in libsosat_b200.so (hand-written CUDA, C-ABI in include/sosat_b200.h);
this package only marshals arguments.  There is no CPU fallback.
"""
from . import _lib, geometry, opt_analyze, ot_geo, ray_trace  # noqa: F401

__version__ = "0.1.0"
