"""Time the fused trace + binning of a frequency sweep: python scripts/prof_sweep.py [n_scan] [n_freq] [grid_n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sosat_optics_b200 import geometry, opt_analyze, ot_geo, ray_trace  # noqa: E402

n_scan = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_freq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
grid_n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
t = ot_geo.SatGeo()
t.n_scan = n_scan
t.y_source = ot_geo.y_lyot + 200
table = np.load(os.path.join(ROOT, "sosat_optics_b200", "data", "fwhm_feedhorns.npy")) if os.path.exists(
    os.path.join(ROOT, "sosat_optics_b200", "data", "fwhm_feedhorns.npy")) else None
freqs = np.linspace(77, 106, n_freq)
if table is not None:
    specs = geometry.freq_specs_from_table(freqs, table)
else:
    specs = [(2 * np.pi / opt_analyze.ghz_to_m(f), 0.5, 0.5) for f in freqs]
planes = None
for nf in (1, n_freq):
    sp = specs[:nf]
    ts = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ray_trace.trace_bin_device(t, [[0, 0, 0]], sp, grid_n, 420.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    best = min(ts[1:])
    print("%s n_scan=%d n_freq=%d grid %d: %.3f ms -> %.3e rays/s, %.3e ray-frequency pairs/s" % (
        os.environ.get("SOSAT_LIB", "default"), n_scan, nf, grid_n, best, n_scan ** 2 / best * 1e3,
        n_scan ** 2 * nf / best * 1e3))
