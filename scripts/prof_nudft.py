"""Time the direct-sum far-field kernel: python scripts/prof_nudft.py [n_pts] [n_ang]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sosat_optics_b200 import _lib  # noqa: E402

lib = _lib.require_gpu()
n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for n_ang in ([int(sys.argv[2])] if len(sys.argv) > 2 else [1024, 65536, 1 << 20]):
    g = torch.Generator(device="cuda").manual_seed(1)
    pts = (torch.rand((n_pts, 4), device="cuda", generator=g) - 0.5) * 0.42
    th = (torch.rand((n_ang, 2), device="cuda", generator=g) - 0.5) * 0.1
    out = torch.empty((n_ang, 2), device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ts = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.sosat_nudft_direct(ctypes.c_void_p(pts.data_ptr()), n_pts, ctypes.c_void_p(th.data_ptr()),
                                          n_ang, 1.0 / 3.33e-3, ctypes.c_void_p(out.data_ptr()), st))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    best = min(ts[1:])
    print("%s n_pts=%d n_ang=%d: %.3f ms -> %.3e pairs/s (%.2f of the 2-MUFU bound 2.327e12)" % (
        os.environ.get("SOSAT_LIB", "default"), n_pts, n_ang, best, n_pts * n_ang / best * 1e3,
        n_pts * n_ang / best * 1e3 / 2.3266e12))
