"""Per-kernel SASS opcode counts of libsosat_b200.so (cuobjdump -sass): the evidence that the
far-field kernels are tcgen05 / TMA / TMEM code and the tracer is fp64 straight-line code.

    python scripts/sass_summary.py [lib.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sosat_optics_b200", "libsosat_b200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.MULTICAST", "UTCBAR", "LDTM", "UTCATOMSWS",
         "UCGABAR", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "MUFU", "FFMA",
         "F2F", "REDG", "ATOMG", "ATOMS", "SHFL", "BAR", "LDC", "LDCU", "LDS", "STS", "LDG", "STG", "STL", "LDL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                       capture_output=True, text=True).stdout.splitlines()
kernels, cur = [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        kernels.append((names[len(kernels)], cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur is not None:
        op = m.group(1)
        parts = op.split(".")
        cur["_total"] += 1
        cur[parts[0]] += 1
        if parts[0] == "MUFU" and len(parts) > 1:
            cur["MUFU." + parts[1]] += 1
        if parts[0] == "UTCHMMA" and "2CTA" in parts:
            cur["UTCHMMA.2CTA"] += 1
        if parts[0] == "UTMALDG" and "MULTICAST" in parts:
            cur["UTMALDG.MULTICAST"] += 1
print(f"# SASS opcode counts per kernel: cuobjdump -sass {os.path.relpath(lib, ROOT)}")
print(f"# {len(kernels)} kernels; static instruction counts (not executed counts)")
tot = collections.Counter()
for name, c in sorted(kernels, key=lambda kc: -kc[1]["_total"]):
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sosat::", "")
    ops = "  ".join(f"{w} {c[w]}" for w in WATCH if c[w])
    print(f"{short:42s} {c['_total']:7d} inst | {ops}")
    tot.update(c)
print("-" * 100)
print(f"{'library total':42s} {tot['_total']:7d} inst | " + "  ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]))
