"""Timing of the batched far-field transform: python scripts/prof_dft_batch.py [n] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sosat_optics_b200 import opt_analyze  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
d = torch.randn(k, n, n, dtype=torch.complex64, device="cuda")
for _ in range(3):
    opt_analyze.far_field_batch(d, device=True)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    opt_analyze.far_field_batch(d, device=True)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
best = min(ts)
print("far_field_batch n=%d batch=%d: best %.4f ms = %.4f ms/transform; %.1f algorithmic TFLOP/s" % (
    n, k, best, best / k, 16.0 * n ** 3 * k / (best * 1e-3) / 1e12))
