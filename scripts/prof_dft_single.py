"""Where a single-shot far-field transform's time goes: Python wrapper vs C-ABI call vs CUDA-graph
replay of the same three launches.   python scripts/prof_dft_single.py [n]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sosat_optics_b200 import _lib, opt_analyze

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _lib.require_gpu()
d = torch.randn(n, n, dtype=torch.complex64, device="cuda")
out = torch.empty_like(d)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = ev(), ev()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


print("python wrapper      best/median ms", timed(lambda: opt_analyze.far_field(d, device=True)))
print("wrapper with out=   best/median ms", timed(lambda: opt_analyze.far_field(d, device=True, out=out)))
ws = opt_analyze._workspace(lib, torch, n)
args = (ctypes.c_void_p(d.data_ptr()), n, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()),
        ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
print("C-ABI call          best/median ms", timed(lambda: lib.sosat_dft2_gemm(*args)))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ws2 = opt_analyze._workspace(lib, torch, n)
    a2 = (args[0], n, args[2], ctypes.c_void_p(ws2.data_ptr()), ws2.numel(), ctypes.c_void_p(s.cuda_stream))
    for _ in range(3):
        lib.sosat_dft2_gemm(*a2)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        lib.sosat_dft2_gemm(*a2)
print("CUDA graph replay   best/median ms", timed(g.replay))
a, b = ev(), ev()
a.record()
for _ in range(50):
    lib.sosat_dft2_gemm(*args)
b.record(); torch.cuda.synchronize()
print("back to back (C)    ms/transform", a.elapsed_time(b) / 50)
