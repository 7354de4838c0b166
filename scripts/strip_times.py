"""Debug: per-strip durations of the fused kernel (library built with -DSOSAT_STRIP_TIMING):
SOSAT_LIB=bin_tmp/lib_striptime.so python scripts/strip_times.py n_scan world"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sosat_optics_b200 import _lib, dist as sdist, opt_analyze, ot_geo, ray_trace
n_scan, world = int(sys.argv[1]), int(sys.argv[2])
t = ot_geo.SatGeo(); t.n_scan = n_scan; t.y_source = ot_geo.y_lyot + 300
t.lambda_ = opt_analyze.ghz_to_m(90); t.k = 2 * np.pi / t.lambda_
spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
rows = torch.from_numpy(sdist.partition_tile_rows(n_scan, world, 0)).to("cuda") if world > 1 else None
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ray_trace.trace_bin_device(t, [[0, 0, 0]], spec, 2048, 420.0, tile_rows=rows)
    b.record(); torch.cuda.synchronize()
    print("ms", a.elapsed_time(b))
lib = ctypes.CDLL(os.environ["SOSAT_LIB"])

tiles_x = (n_scan + 15) // 16
n_rows = len(rows) if rows is not None else tiles_x
n_tiles = tiles_x * n_rows
sl = max(1, min(64, n_tiles // (8 * 592)))
strips_x = (tiles_x + sl - 1) // sl
n = min(strips_x * n_rows, 1 << 17)
t0 = np.zeros(n, np.int64); t1 = np.zeros(n, np.int64); sm = np.zeros(n, np.int32)
lib.sosat_debug_strip_times(t0.ctypes.data_as(ctypes.c_void_p), t1.ctypes.data_as(ctypes.c_void_p), sm.ctypes.data_as(ctypes.c_void_p), n)
d = (t1 - t0) / 1e3
start = t0.min(); end = t1.max()
print("strips", n, "strip_len", sl, "kernel span us", (end - start) / 1e3)
print("strip duration us: mean %.1f median %.1f p90 %.1f p99 %.1f max %.1f" % (d.mean(), np.median(d), np.percentile(d, 90), np.percentile(d, 99), d.max()))
# idle at the end per SM: last strip end per SM
last = np.array([t1[sm == s].max() for s in np.unique(sm)])
print("SMs", len(last), "finish spread us: mean idle %.1f max idle %.1f" % (((end - last) / 1e3).mean(), ((end - last) / 1e3).max()))
first = np.array([t0[sm == s].min() for s in np.unique(sm)])
print("start spread us: mean %.1f max %.1f" % (((first - start) / 1e3).mean(), ((first - start) / 1e3).max()))
# by column position
sx = np.arange(n) % strips_x
for k in range(0, strips_x, max(1, strips_x // 16)):
    print("sx", k, "mean dur us %.1f" % d[sx == k].mean())
busy = d.sum() / (592 * (end - start) / 1e3)
print("CTA busy fraction %.3f" % busy)
