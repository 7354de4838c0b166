"""Executed warp instructions per CUDA source line from an ncu report captured with
--import-source on (kernel built with -lineinfo):
    python scripts/ncu_lines.py report.ncu-rep [rays]  -> per-line counts, largest first"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cnt, samp, text = collections.Counter(), collections.Counter(), {}
path = None
rows = None
for line in txt.splitlines():
    if line.startswith('"File Path"'):
        path = next(csv.reader([line]))[1].split("/")[-1]
        continue
    if line.startswith('"Line No"'):
        hdr = next(csv.reader([line]))
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if not line.startswith('"') or path is None:
        continue
    r = next(csv.reader([line]))
    if r[0] == "" or not r[0].isdigit():
        continue
    key = (path, int(r[0]))
    num = lambda v: int(v) if v.isdigit() else 0  # noqa: E731
    cnt[key] += num(r[i_inst])
    samp[key] += num(r[i_samp])
    text[key] = r[1].strip()[:90]
tot = sum(cnt.values())
div = rays / 32 if rays else 1.0
print("total warp instructions %d (%.1f per ray)" % (tot, tot / div))
for key, n in cnt.most_common(70):
    print("%-18s %5d  %8.2f  %5.1f%%  smp %6d  %s" % (key[0], key[1], n / div, 100.0 * n / tot, samp[key], text[key]))
