"""Timing of the per-ray output modes of the tracer: python scripts/prof_rows.py [n_scan]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sosat_optics_b200 import _lib, geometry, opt_analyze, ot_geo, ray_trace  # noqa: E402

n_scan = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.require_gpu()
t = ot_geo.SatGeo()
t.n_scan = n_scan
t.y_source = ot_geo.y_lyot + 300
t.lambda_ = opt_analyze.ghz_to_m(90)
t.k = 2 * np.pi / t.lambda_
ray_trace._upload_geometry(lib, torch, t)
freq = geometry.make_freq(t)
rx = np.zeros(3)
n = n_scan * n_scan
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
stats = torch.zeros(8, dtype=torch.float64, device="cuda")
xyb = torch.empty((n, 4), dtype=torch.float32, device="cuda")
out = torch.empty((17, n), dtype=torch.float64, device="cuda")
dp = rx.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


ms = timed(lambda: _lib.check(lib.sosat_trace_samples(dp, ctypes.byref(freq), n_scan, 0, n, 0.0,
                                                      ctypes.c_void_p(xyb.data_ptr()),
                                                      ctypes.c_void_p(stats.data_ptr()), st)))
print("compact float4 samples n_scan=%d: %.3f ms -> %.3e rays/s, %.0f GB/s written" % (
    n_scan, ms, n / ms * 1e3, n * 16 / ms / 1e6))
ms = timed(lambda: _lib.check(lib.sosat_trace_rays(dp, ctypes.byref(freq), n_scan, 0, n,
                                                   ctypes.c_void_p(out.data_ptr()),
                                                   ctypes.c_void_p(stats.data_ptr()), st)))
print("reference table out[17][n] n_scan=%d: %.3f ms -> %.3e rays/s, %.0f GB/s written" % (
    n_scan, ms, n / ms * 1e3, n * 136 / ms / 1e6))
