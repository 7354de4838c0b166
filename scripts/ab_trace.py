"""A/B timing of the fused trace+bin kernel at the benchmark size under different env settings
(one process, one gpurun call):  python scripts/ab_trace.py  [n_scan]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sosat_optics_b200 import _lib, opt_analyze, ot_geo, ray_trace

n_scan = int(sys.argv[1]) if len(sys.argv) > 1 else 31623
t = ot_geo.SatGeo()
t.n_scan = n_scan
t.y_source = ot_geo.y_lyot + 300
t.lambda_ = opt_analyze.ghz_to_m(90)
t.k = 2 * np.pi / t.lambda_
spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
re = torch.zeros((1, 1, 2048, 2048), device="cuda")
im, cnt = torch.zeros_like(re), torch.zeros((1, 2048, 2048), device="cuda")
st = torch.zeros((1, 8), dtype=torch.float64, device="cuda")

VARIANTS = [{"SOSAT_TRACE_PREDICT": "0"}, {"SOSAT_TRACE_PREDICT": "1"},
            {"SOSAT_TRACE_PREDICT": "1", "SOSAT_STRIP_LEN": "32"},
            {"SOSAT_TRACE_PREDICT": "1", "SOSAT_STRIP_LEN": "64"}]
ROWS = None
if len(sys.argv) > 2:   # emulate rank 0 of a W-rank cyclic tile-row deal
    from sosat_optics_b200 import dist as sdist
    ROWS = torch.from_numpy(sdist.partition_tile_rows(n_scan, int(sys.argv[2]), 0)).to("cuda")
base = None
for env in VARIANTS:
    for k in ("SOSAT_TRACE_PREDICT", "SOSAT_STRIP_LEN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for rep in range(4):
        for x in (re, im, cnt, st):
            x.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ray_trace.trace_bin_device(t, [[0, 0, 0]], spec, 2048, 420.0, planes=(re, im, cnt), stats=st,
                                   tile_rows=ROWS)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    s = st.cpu().numpy()[0]
    if base is None:
        base = (s.copy(), cnt.clone(), re.clone())
    dc = float((cnt - base[1]).abs().sum().item())
    dr = float((re - base[2]).abs().max().item() / base[2].abs().max().item())
    print(env, "ms best %.3f median %.3f" % (min(ts[1:]), sorted(ts[1:])[1]),
          "rays/s %.4g" % (n_scan * n_scan / (min(ts[1:]) * 1e-3)),
          "n_valid %d n_binned %d n_slow %d" % (s[0], s[7], s[5]),
          "dcnt %g dre %.2e dsumopl %.3e" % (dc, dr, s[1] - base[0][1]), flush=True)
