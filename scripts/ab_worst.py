"""Worst rays of a schedule against the 0+5 fp64 solve.  python scripts/ab_worst.py 7 0 0 170"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sosat_optics_b200 import ot_geo, ray_trace  # noqa: E402

var = sys.argv[1]
rx = [float(v) for v in sys.argv[2:5]]
t = ot_geo.SatGeo()
t.n_scan = 1000
t.y_source = ot_geo.y_lyot + 300
os.environ["SOSAT_TRACE_VARIANT"] = "0"
ref = ray_trace.rx_to_lyot(rx, t)
os.environ["SOSAT_TRACE_VARIANT"] = var
out = ray_trace.rx_to_lyot(rx, t)
m = (ref[4] != 0) & (out[4] != 0)
d = np.abs(out[3] - ref[3]) * m
print("hist of log10 |dOPL| (valid rays):", np.histogram(np.log10(d[m] + 1e-16), bins=np.arange(-16, -4))[0])
idx = np.argsort(-d)[:12]
for i in idx:
    print("ray %7d (iphi %4d ith %4d) dOPL %.2e  ap (%.2f, %.2f) r %.1f  ddir %.1e  dpos %.1e" % (
        i, i // 1000, i % 1000, d[i], ref[0][i], ref[2][i], np.hypot(ref[0][i], ref[2][i]),
        max(abs(out[r][i] - ref[r][i]) for r in (8, 9, 10)), max(abs(out[r][i] - ref[r][i]) for r in (0, 2))))
