"""Profiling driver for the far-field transform: python scripts/prof_dft.py [n] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sosat_optics_b200 import opt_analyze  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = torch.randn(n, n, dtype=torch.complex64, device="cuda")
for _ in range(3):
    opt_analyze.far_field(d, device=True)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    opt_analyze.far_field(d, device=True)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
# back-to-back throughput (host launch overhead included)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    opt_analyze.far_field(d, device=True)
b.record()
torch.cuda.synchronize()
print("far_field n=%d: best %.4f ms median %.4f ms; back-to-back %.4f ms/transform; %.1f algorithmic TFLOP/s" % (
    n, ts[0], ts[len(ts) // 2], a.elapsed_time(b) / reps, 16.0 * n ** 3 / (ts[0] * 1e-3) / 1e12))
