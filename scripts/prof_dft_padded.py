"""Zero-padded batch far field (compact K) against explicit padding: python scripts/prof_dft_padded.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sosat_optics_b200 import opt_analyze as oa

ev = lambda: torch.cuda.Event(enable_timing=True)
for k, n_in, pad in ((16, 512, 256), (64, 512, 256), (276, 512, 256), (64, 256, 128), (16, 1024, 512)):
    d = torch.randn(k, n_in, n_in, dtype=torch.complex64, device="cuda")
    res = {}
    for name, fn in (("compact", lambda: oa.far_field_batch(d, device=True, pad=pad)),
                     ("explicit", lambda: oa.far_field_batch(torch.nn.functional.pad(d, (pad,) * 4), device=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            a, b = ev(), ev()
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name] = min(ts)
    n = n_in + 2 * pad
    print("K=%3d  %4d^2 -> %4d^2: compact %.3f ms (%.4f per field)  explicit padding %.3f ms (%.4f per field)  x%.2f"
          % (k, n_in, n, res["compact"], res["compact"] / k, res["explicit"], res["explicit"] / k,
             res["explicit"] / res["compact"]), flush=True)
    oa.release_workspaces()
