"""A/B of a Newton schedule (SOSAT_TRACE_VARIANT) against the 0+5 fp64 schedule: row differences
on 10^6-ray bundles for several feeds, then the fused kernel's time.  python scripts/ab_variant.py 7"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sosat_optics_b200 import ot_geo, ray_trace  # noqa: E402

var = sys.argv[1] if len(sys.argv) > 1 else "7"
t = ot_geo.SatGeo()
t.n_scan = 1000
t.y_source = ot_geo.y_lyot + 300
for rx in ([0, 0, 0], [150, 0, 0], [-120, 0, 80], [0, 0, 170], [60, 0, -90]):
    os.environ["SOSAT_TRACE_VARIANT"] = "0"
    ref = ray_trace.rx_to_lyot(rx, t)
    os.environ["SOSAT_TRACE_VARIANT"] = var
    out = ray_trace.rx_to_lyot(rx, t)
    vr, vo = ref[4] != 0, out[4] != 0
    m = vr & vo
    d = [np.abs(out[r][m] - ref[r][m]) for r in range(11)]
    tame = m & (np.hypot(ref[0], ref[2]) < 400.0)
    print("rx", rx, "mask diff", int((vr != vo).sum()), "valid", int(vr.sum()),
          "opl tame %.2e" % np.abs(out[3][tame] - ref[3][tame]).max(),
          "max %.2e med %.2e  pos %.2e %.2e  nrm %.2e dir %.2e amp %.2e" % (
              d[3].max(), np.median(d[3]), d[0].max(), d[2].max(), max(x.max() for x in d[5:8]),
              max(x.max() for x in d[8:11]), np.abs(out[4][m] / ref[4][m] - 1).max()), flush=True)
for rxs in (("150,0,0", "-120,0,80", "0,0,170") if os.environ.get("AB_TIME", "1") == "1" else ()):
    for v in ("6", var):
        env = dict(os.environ, SOSAT_TRACE_VARIANT=v, PROF_REPS="6", PROF_RX=rxs)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/prof_trace.py"), "8192"], env=env,
                           capture_output=True, text=True)
        print("rx", rxs, "variant", v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
for n in (("8192", "31623") if os.environ.get("AB_TIME", "1") == "1" else ()):
    for v in ("6", var, "6", var):
        env = dict(os.environ, SOSAT_TRACE_VARIANT=v, PROF_REPS="6")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/prof_trace.py"), n], env=env,
                           capture_output=True, text=True)
        print("variant", v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
