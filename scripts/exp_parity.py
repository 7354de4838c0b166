"""Experiment helper: parity (worst and median ray) + speed of the library named by $SOSAT_LIB."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import golden, tele_geo_from_golden
from sosat_optics_b200 import ray_trace
from oracle import oracle
g = golden("trace_cfg1_n60_106ghz.npz"); t = tele_geo_from_golden(g); ref = g["out"]
out = ray_trace.rx_to_lyot([0, 0, 0], t)
m = ref[4] != 0
d = np.abs(out[3][m] - ref[3][m])
wild = np.hypot(ref[0][m], ref[2][m]) > 400
print(os.environ.get("SOSAT_LIB", "default"), "mask", bool(((out[4] != 0) == m).all()),
      "dOPL max %.2e median %.2e max(non-wild) %.2e  dpos max %.2e" % (d.max(), np.median(d), d[~wild].max(), np.abs(out[0][m]-ref[0][m]).max()))
t.n_scan = 1000
for rx in ([0, 0, 0], [0, 0, 150]):
    o = ray_trace.rx_to_lyot(rx, t); r = oracle.rx_to_lyot(rx, t); m = r[4] != 0
    d = np.abs(o[3][m] - r[3][m]); wild = np.hypot(r[0][m], r[2][m]) > 400
    print("   1e6 rays rx", rx, "mask mism", int(((o[4] != 0) != m).sum()), "dOPL max %.2e p99.9 %.2e median %.2e max(non-wild) %.2e" % (d.max(), np.quantile(d, 0.999), np.median(d), d[~wild].max()))
t.n_scan = 8192
spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
planes = None; best = 1e9
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a, b, c, s = ray_trace.trace_bin_device(t, [[0, 0, 0]], spec, 2048, 420.0, planes=planes); e1.record()
    torch.cuda.synchronize(); planes = (a, b, c)
    if rep: best = min(best, e0.elapsed_time(e1))
print("   bin 8192^2: %.2f ms %.3e rays/s" % (best, 8192 ** 2 / best * 1e3))
