"""Profiling driver: a few launches of the fused trace+bin kernel (n_scan rays per side,
2048^2 bins) for ncu.  python scripts/prof_trace.py [n_scan]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sosat_optics_b200 import opt_analyze, ot_geo, ray_trace  # noqa: E402

n_scan = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
t = ot_geo.SatGeo()
t.n_scan = n_scan
t.y_source = ot_geo.y_lyot + 300
t.lambda_ = opt_analyze.ghz_to_m(90)
t.k = 2 * np.pi / t.lambda_
spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
planes = None
reps = int(os.environ.get("PROF_REPS", "3"))
rx = [[float(v) for v in os.environ.get("PROF_RX", "0,0,0").split(",")]]
ts = []
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    re, im, cnt, st = ray_trace.trace_bin_device(t, rx, spec, 2048, 420.0, planes=planes)
    e1.record()
    torch.cuda.synchronize()
    planes = (re, im, cnt)
    ts.append(e0.elapsed_time(e1))
    if reps <= 3:
        print("rep", rep, "%.3f ms  %.3e rays/s" % (ts[-1], n_scan ** 2 / ts[-1] * 1e3))
ts = sorted(ts[1:])
print("%s n_scan=%d: best %.3f ms median %.3f ms -> %.3e rays/s (best)" % (
    os.environ.get("SOSAT_LIB", "default"), n_scan, ts[0], ts[len(ts) // 2], n_scan ** 2 / ts[0] * 1e3))
