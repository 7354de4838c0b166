"""Dynamic opcode mix and stall samples per opcode from an ncu report's SASS source page.

    python scripts/ncu_opmix.py report.ncu-rep [rays]   # rays: divide counts to get per-ray figures
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
cnt, samp, nis = collections.Counter(), collections.Counter(), collections.Counter()
stall_tot = collections.Counter()
rows = []
for r in rd:
    src = r["Source"].strip()
    m = re.match(r"(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", src)
    if not m:
        continue
    op = m.group(1)
    if op in ("MUFU", "F2F", "I2F", "F2I", "FRND"):
        op += m.group(2)
    n = int(r["Instructions Executed"])
    cnt[op] += n
    samp[op] += int(r["Warp Stall Sampling (All Samples)"])
    nis[op] += int(r["Warp Stall Sampling (Not-issued Samples)"])
    for k, v in r.items():
        if k.startswith("stall_") and "(Not Issued)" in k and v and v != "0":
            stall_tot[k.replace(" (Not Issued)", "")] += int(v)
    rows.append((n, src))
tot = sum(cnt.values())
stot = sum(samp.values())
div = (rays / 32.0) if rays else 1.0
print(f"warp instructions executed: {tot}" + (f"  = {tot / div:.1f} per ray" if rays else ""))
print(f"{'opcode':22s} {'warp-inst':>12s} {'per ray':>9s} {'share':>7s} {'samples':>8s} {'smp share':>9s}")
for op, n in cnt.most_common(40):
    print(f"{op:22s} {n:12d} {n / div:9.1f} {100 * n / tot:6.1f}% {samp[op]:8d} {100 * samp[op] / max(stot, 1):8.1f}%")
print("not-issued stall samples by reason:")
st = sum(stall_tot.values())
for k, v in stall_tot.most_common(12):
    print(f"  {k:28s} {v:8d} {100 * v / max(st, 1):6.1f}%")
