"""Build experimental variants of libsosat_b200.so with extra nvcc flags into bin_tmp/ for A/B
timing on the GPU box:   python scripts/ab_build.py name "-DFOO=1 -DBAR=2" [name2 "flags2" ...]
(AB_SOURCES=dft_gemm.cu,trace.cu selects which sources are recompiled with the flags; default trace.cu)
Run with SOSAT_LIB=bin_tmp/lib_<name>.so python scripts/prof_trace.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sosat_optics_b200 import build as B  # noqa: E402

out_dir = os.path.join(ROOT, "bin_tmp")
os.makedirs(out_dir, exist_ok=True)
args = sys.argv[1:]
procs = []
for name, flags in zip(args[0::2], args[1::2]):
    objs = []
    for src in B.SOURCES:
        o = os.path.join(out_dir, f"{name}_{src.replace('.cu', '.o')}")
        if src not in os.environ.get("AB_SOURCES", "trace.cu").split(",") and os.path.exists(os.path.join(B.CSRC, "_obj", src.replace(".cu", ".o"))):
            objs.append(os.path.join(B.CSRC, "_obj", src.replace(".cu", ".o")))
            continue
        objs.append(o)
        cmd = [B._nvcc(), *B.NVCC_FLAGS, *flags.split(), "-I", B.INCLUDE, "-I", B.CSRC, "-c",
               os.path.join(B.CSRC, src), "-o", o]
        procs.append((name, subprocess.Popen(cmd)))
    globals().setdefault("links", []).append((name, objs))
for name, p in procs:
    assert p.wait() == 0, name
for name, objs in links:
    lib = os.path.join(out_dir, f"lib_{name}.so")
    subprocess.check_call([B._nvcc(), "-shared", "-o", lib, *objs, "-cudart", "static", "-gencode",
                           "arch=compute_100a,code=sm_100a"])
    print(lib)
