"""Build libsosat_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m sosat_optics_b200.build [--force] [--verbose]

The shared library exposes only the C-ABI declared in include/sosat_b200.h; Python binds it
with ctypes (sosat_optics_b200/_lib.py).  cudart is linked statically, so the library loads
(and exports its symbols) on a machine without a GPU or driver.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG, "libsosat_b200.so")
SOURCES = ["capi.cu", "trace.cu", "dft_gemm.cu", "nudft.cu", "reduce.cu", "peer.cu", "gridder.cu"]
HEADERS = ["common.cuh", "trace_device.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """Compile the product library; ``experiments=True`` builds ``libsosat_b200_exp.so`` instead,
    with -DSOSAT_EXPERIMENTS (the kernels that lost their A/B runs, for scripts/ab_* and the
    variant-agreement tests: point SOSAT_LIB at it)."""
    if experiments:
        return _build("libsosat_b200_exp.so", "_obj_exp", ["-DSOSAT_EXPERIMENTS"], force, verbose)
    return _build("libsosat_b200.so", "_obj", [], force, verbose)


def _build(lib_name: str, obj_name: str, extra, force: bool, verbose: bool) -> str:
    LIB_PATH = os.path.join(PKG, lib_name)  # noqa: N806
    deps = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(INCLUDE, "sosat_b200.h"),
                                                       os.path.abspath(__file__)]
    objs = []
    obj_dir = os.path.join(CSRC, obj_name)
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, deps + [s]):
            cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", INCLUDE, "-I", CSRC, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                experiments="--experiments" in sys.argv))
