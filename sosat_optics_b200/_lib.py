"""ctypes binding of libsosat_b200.so (the C-ABI declared in include/sosat_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible
when a compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t,
                    c_uint8, c_void_p)

NUM_SURFACES = 6
OUT_ROWS = 17
MAX_FREQS = 64
NUM_STATS = 8
PEER_HANDLE_BYTES = 64  # SOSAT_PEER_HANDLE_BYTES
DFT_FORWARD, DFT_INVERSE, DFT_B2A = 0, 1, 2
(STAT_N_VALID, STAT_SUM_OPL, STAT_SUM_DX, STAT_SUM_DY, STAT_SUM_DZ, STAT_N_SLOW,
 STAT_N_BRACKET_FAIL, STAT_N_BINNED) = range(8)

LIB_PATH = os.environ.get("SOSAT_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "libsosat_b200.so")


class Surface(Structure):
    _fields_ = [(n, c_double) for n in
                ("c", "k", "a1", "a2", "a3", "y0", "n_in", "n_out", "t_lo", "t_hi")]


class Geom(Structure):
    _fields_ = [("surf", Surface * NUM_SURFACES)] + [
        (n, c_double) for n in ("y_lyot", "y_source", "clip_l3", "clip_stop", "cone")]


class Freq(Structure):
    _fields_ = [("fwhp_x", c_double), ("fwhp_y", c_double), ("k_over_2e3", c_double)]


# name -> (restype, argtypes); mirrors include/sosat_b200.h one to one
SIGNATURES = {
    "sosat_abi_version": (c_int, []),
    "sosat_build_flags": (c_int, []),
    "sosat_last_error": (c_char_p, []),
    "sosat_device_count": (c_int, []),
    "sosat_set_device": (c_int, [c_int]),
    "sosat_device_info": (c_int, [c_char_p, POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "sosat_set_geometry": (c_int, [POINTER(Geom), c_void_p]),
    "sosat_trace_rays": (c_int, [POINTER(c_double), POINTER(Freq), c_int, c_int64, c_int64,
                                 c_void_p, c_void_p, c_void_p]),
    "sosat_trace_samples": (c_int, [POINTER(c_double), POINTER(Freq), c_int, c_int64, c_int64,
                                    c_double, c_void_p, c_void_p, c_void_p]),
    "sosat_rx_to_lyot_host": (c_int, [POINTER(Geom), POINTER(c_double), POINTER(Freq), c_int,
                                      c_int64, c_int64, POINTER(c_double), POINTER(c_double)]),
    "sosat_near_field_arrays": (c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "sosat_trace_bin": (c_int, [c_void_p, c_int, POINTER(Freq), c_int, c_int, c_int64, c_int64,
                                c_int, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "sosat_trace_bin_tiles": (c_int, [c_void_p, c_int, POINTER(Freq), c_int, c_int, c_void_p, c_int,
                                      c_int, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "sosat_bins_to_field": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p,
                                    c_void_p]),
    "sosat_bins_to_field_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_double,
                                          c_double, c_void_p, c_void_p]),
    "sosat_bins_to_field_stats_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                                c_void_p, POINTER(c_double), c_double, c_void_p,
                                                c_void_p]),
    "sosat_dft2_workspace_bytes": (c_size_t, [c_int]),
    "sosat_dft2_forget": (c_int, [c_void_p]),
    "sosat_dft2_gemm": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sosat_dft2_gemm_kind": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t,
                                     c_void_p]),
    "sosat_dft2_batch_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sosat_dft2_gemm_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t,
                                      c_void_p]),
    "sosat_dft2_batch_padded_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sosat_dft2_gemm_batch_padded": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                             c_size_t, c_void_p]),
    "sosat_dft2_zoom": (c_int, [c_void_p, c_int, c_void_p, c_int, POINTER(c_double),
                                POINTER(c_double), POINTER(c_double), POINTER(c_double), c_double,
                                c_void_p, c_size_t, c_void_p]),
    "sosat_b2a": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                          c_void_p]),
    "sosat_horn_aperture": (c_int, [c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p,
                                    c_void_p]),
    "sosat_cmul": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "sosat_a2b": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                          c_void_p]),
    "sosat_a2b_host": (c_int, [POINTER(c_double), POINTER(c_double), c_int, POINTER(c_double),
                               POINTER(c_double)]),
    "sosat_nudft_direct": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_double, c_void_p,
                                   c_void_p]),
    "sosat_near_field_metrics": (c_int, [c_void_p, c_int64, c_int, c_int64, c_double, POINTER(Freq),
                                         c_int, c_void_p, c_void_p, c_void_p]),
    "sosat_ff_fwhm_indices": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sosat_ff_fwhm_indices_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                            c_void_p]),
    "sosat_bins_reduce_fields": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_int64,
                                         c_int64, c_int, c_int, c_int64, c_int64, POINTER(c_double),
                                         c_double, POINTER(c_void_p), c_int, c_int64, c_void_p,
                                         c_void_p]),
    "sosat_sim2d_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "sosat_sim2d_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                    c_double, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "sosat_sim2d_eval": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                 c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    "sosat_sim_data_work_doubles": (c_size_t, [c_int, c_int]),
    "sosat_sim_data": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "sosat_peer_barrier": (c_int, [POINTER(c_void_p), c_int, c_int, c_int64, ctypes.c_uint64, c_double,
                                   c_void_p]),
    "sosat_measure_fp64_peak": (c_int, [POINTER(c_double), POINTER(c_double)]),
    "sosat_peer_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "sosat_peer_free": (c_int, [c_void_p]),
    "sosat_peer_export": (c_int, [c_void_p, c_void_p]),
    "sosat_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "sosat_peer_close": (c_int, [c_void_p]),
    "sosat_bins_reduce_field": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_int64,
                                        c_int64, c_int64, c_double, c_double, POINTER(c_void_p),
                                        c_int, c_int64, c_void_p, c_void_p]),
}

_lib = None


class SosatError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m sosat_optics_b200.build` "
            "(needs nvcc).  sosat_optics_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.sosat_abi_version() != 1:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.sosat_abi_version()} != 1")
    _lib = lib
    return lib


def has_experiments() -> bool:
    """True if the loaded library was built with -DSOSAT_EXPERIMENTS (scripts/ab_build.py)."""
    return bool(load().sosat_build_flags() & 1)


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sosat_last_error()
        raise SosatError(f"libsosat_b200 error {rc}: {msg.decode() if msg else ''}")


_gpu_ok = False


def require_gpu():
    """The product path needs a B200; say so loudly instead of computing on the CPU."""
    global _gpu_ok
    lib = load()
    if not _gpu_ok:   # checked until it succeeds once
        if lib.sosat_device_count() <= 0:
            raise SosatError("no CUDA device visible: sosat_optics_b200 runs only on the GPU "
                             "(there is deliberately no CPU fallback)")
        _gpu_ok = True
    return lib


PINNED_LIMIT_BYTES = 1 << 30
STAGE_BYTES = 64 << 20


def to_numpy(t):
    """Device tensor -> numpy array through PINNED host memory (one DMA at PCIe speed, no
    pageable staging and no page faults on a fresh allocation; torch's caching host allocator
    recycles the block when the array is dropped).  The array is a view of the pinned tensor,
    which it keeps alive.  Results above PINNED_LIMIT_BYTES (e.g. rx_to_lyot's 136 B/ray table at
    n_scan = 4096: 2.3 GB) go through two pinned staging buffers of STAGE_BYTES: the DMA of chunk
    k + 1 overlaps the host copy of chunk k into the pageable result (4-5x a pageable cudaMemcpy,
    without page-locking gigabytes)."""
    import numpy as np
    import torch
    if t.device.type != "cuda":
        return t.numpy()
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return t.cpu().numpy()
    if nbytes > PINNED_LIMIT_BYTES:
        try:
            return _to_numpy_staged(t, np, torch)
        except RuntimeError:
            return t.cpu().numpy()
    try:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:  # page-locked memory exhausted: the pageable copy still works
        return t.cpu().numpy()
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def _to_numpy_staged(t, np, torch):
    src = t.contiguous().view(-1).view(torch.uint8)
    n = src.numel()
    out = torch.empty(n, dtype=torch.uint8)  # pageable result
    stage = [torch.empty(STAGE_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    stream = torch.cuda.current_stream()
    offs = list(range(0, n, STAGE_BYTES))
    for i, o in enumerate(offs + [None]):
        if o is not None:
            m = min(STAGE_BYTES, n - o)
            stage[i & 1][:m].copy_(src[o:o + m], non_blocking=True)
            done[i & 1].record(stream)
        if i > 0:  # drain the previous chunk while this one is in flight
            po = offs[i - 1]
            pm = min(STAGE_BYTES, n - po)
            done[(i - 1) & 1].synchronize()
            out[po:po + pm].copy_(stage[(i - 1) & 1][:pm])
    return out.view(t.dtype).view(t.shape).numpy()
