"""Drop-in for the hot-path functions of the reference's ``sosat_optics/opt_analyze.py``.

======================  =======================  ==============================================
function                reference                device code
======================  =======================  ==============================================
``a2b``                 opt_analyze.py:220-230   ``sosat_a2b``: tcgen05 bf16x3 DFT-GEMM (dft_gemm.cu)
``b2a``                 opt_analyze.py:208-217   ``sosat_b2a``: same GEMM path, conjugate twiddles
``beam_convolve``       opt_analyze.py:60-88     ``sosat_horn_aperture`` + 3 x ``sosat_dft2_gemm_kind``
                                                 + ``sosat_cmul``
``far_field``           -- (north_star)          ``sosat_dft2_gemm`` on complex64 device/host data
``far_field_irregular`` -- (north_star)          ``sosat_nudft_direct`` direct sum (nudft.cu)
``zero_pad``            opt_analyze.py:20-32     host index bookkeeping (np.pad + coordinates)
``coords_spat_to_ang``  opt_analyze.py:274-299   host label arithmetic on N values
``coords_ang_to_spat``  opt_analyze.py:233-271   host label arithmetic on N values
``ghz_to_m, m_to_ghz``  opt_analyze.py:42-57     scalar helpers
======================  =======================  ==============================================
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np

from . import _lib
from .ray_trace import _stream, _torch

__all__ = ["a2b", "b2a", "beam_convolve", "far_field", "far_field_batch", "far_field_irregular",
           "zero_pad",
           "coords_spat_to_ang", "coords_ang_to_spat", "ghz_to_m", "m_to_ghz", "rad_to_arcmin"]


def rad_to_arcmin(rads):
    return rads * 180 * 60 / np.pi


def ghz_to_m(freq_ghz):
    """Frequency [GHz] -> wavelength [m] with c = 3e8 exactly (opt_analyze.py:42-48)."""
    ff_hz = freq_ghz * 1e9
    return (3 * 10 ** 8) / ff_hz


def m_to_ghz(wavelength):
    ff_hz = (3 * 10 ** 8) / wavelength
    return ff_hz / 1e9


def zero_pad(x_in, y_in, beam_in, pts):
    """Pad ``beam_in`` by ``pts`` zeros on every side, rebuild the coordinate grids with the
    reference's diagonal spacing ``|x[0,0]-x[1,1]|``, map NaN -> 0 and return the TRANSPOSED
    field -- all quirks of opt_analyze.py:20-32 kept."""
    x_int = abs(x_in[0, 0] - x_in[1, 1])
    y_int = abs(y_in[0, 0] - y_in[1, 1])
    beam_out = np.pad(beam_in, pts, mode="constant")
    x_new = np.array(np.arange(len(beam_out))) * x_int
    y_new = np.array(np.arange(len(beam_out))) * y_int
    x_new -= np.mean(x_new)
    y_new -= np.mean(y_new)
    x_new, y_new = np.meshgrid(x_new, y_new)
    beam_out = np.where(np.isnan(beam_out) == True, 0, beam_out)  # noqa: E712
    return x_new + np.mean(x_in), y_new + np.mean(y_in), beam_out.T


def coords_spat_to_ang(x, y, freq):
    """Angle labels of the far-field grid (opt_analyze.py:274-299).  Kept bit-for-bit,
    including ``len(x)`` for both axes and the N/(N-1) stretch of the linspace labels."""
    ff_ghz = freq * 1e9
    lam = (3 * 10 ** 8) / ff_ghz
    delta_x = abs(np.max(x) - np.min(x)) / (len(x) - 1)
    delta_y = abs(np.max(y) - np.min(y)) / (len(y) - 1)
    x_len = len(x)
    y_len = len(y)
    alpha = lam / delta_x
    beta = lam / delta_y
    delta_th = alpha / x_len
    delta_ph = beta / y_len
    x_ang = np.linspace(-int((len(x) / 2)), int((len(x) / 2)), int((len(x)))) * delta_th
    y_ang = np.linspace(-int((len(x) / 2)), int((len(x) / 2)), int((len(x)))) * delta_ph
    x_ang, y_ang = np.meshgrid(x_ang, y_ang)
    return x_ang, y_ang


def coords_ang_to_spat(theta_x, theta_y, freq):
    """Spatial labels [cm] of the aperture grid (opt_analyze.py:233-271).  Kept as written:
    the second ``delta_th`` assignment (from ``theta_y``) overwrites the first, so both axes use
    the theta_y increment, and ``len(theta_x)`` sizes both axes."""
    ff_ghz = freq * 1e9
    lam = (3 * 10 ** 8) / ff_ghz
    delta_th = abs(np.max(theta_x) - np.min(theta_x)) / (len(theta_x) - 1)
    delta_th = abs(np.max(theta_y) - np.min(theta_y)) / (len(theta_y) - 1)
    x_len = len(theta_x)
    y_len = len(theta_y)
    alpha = lam / delta_th
    beta = lam / delta_th
    delta_x = alpha / x_len
    delta_y = beta / y_len
    half, cnt = int(len(theta_x) / 2), int(len(theta_x))
    x_spat = np.linspace(-half, half, cnt) * delta_x * 1e2
    y_spat = np.linspace(-half, half, cnt) * delta_y * 1e2
    x_spat, y_spat = np.meshgrid(x_spat, y_spat)
    return x_spat, y_spat


# --- DFT-GEMM workspaces: one per (device, stream, n, transform kind), kept alive so the twiddles
# of each kind are generated once.  The stream is part of the key because a workspace holds the
# scratch of a transform in flight (b1 / d1 / c64) and its twiddles are written by a kernel on
# the stream of the first call: two streams (or threads with their own streams) transforming the
# same size must not share one. -----------------------------------------------------------------
_workspaces = {}


def _workspace(lib, torch, n: int, kind: int = _lib.DFT_FORWARD):
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream, int(n), int(kind))
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(lib.sosat_dft2_workspace_bytes(int(n)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        ptr = ws.data_ptr()
        weakref.finalize(ws, lambda p=ptr: lib.sosat_dft2_forget(ctypes.c_void_p(p)))
        _workspaces[key] = ws
    return ws


def release_workspaces():
    _workspaces.clear()
    _batch_workspaces.clear()


def a2b(apert, phi):
    """FFT aperture plane to angular space (opt_analyze.py:220-230).

    ``apert``: real [N, N] amplitude (``abs`` is taken), ``phi``: phase in DEGREES.  Returns
    ``(beam_complex complex128 [N, N], phase float64 [N, N])`` with
    ``beam_complex = fftshift(fft2(fftshift(|A| exp(i phi pi/180))))`` computed as the dense
    separable contraction W b W^T on tensor cores, and ``phase = arctan2(imag, real)``.

    Precision: the contraction runs as a 3-term split-bf16 product with fp32 accumulation, so the
    complex128 array carries about fp32 information -- measured <= 1.5e-5 of the peak |B| on
    worst-case random fields at N = 1024 and ~5e-6 on the reference's own fields, against the 1e-4
    of peak that the hot path is held to.  ``phase`` is the arctan2 of those values: accurate
    where |B| is well above that error floor (5e-3 rad at 1e-2 of peak), meaningless -- as in any
    floating-point FFT -- where |B| is at the floor."""
    return _amp_phase_transform("a2b", apert, phi, _lib.DFT_FORWARD)


def _amp_phase_transform(fn_name, amp, phase_deg, kind):
    lib = _lib.require_gpu()
    torch = _torch()
    amp = np.asarray(amp, dtype=np.float64)
    phase_deg = np.asarray(phase_deg, dtype=np.float64)
    if amp.ndim != 2 or amp.shape[0] != amp.shape[1] or amp.shape != phase_deg.shape:
        raise ValueError(f"{fn_name} expects square [N, N] amplitude and phase arrays of equal shape")
    n = amp.shape[0]
    d_a = torch.from_numpy(np.ascontiguousarray(amp)).to("cuda")
    d_p = torch.from_numpy(np.ascontiguousarray(phase_deg)).to("cuda")
    field = torch.empty((n, n, 2), dtype=torch.float64, device="cuda")
    phase = torch.empty((n, n), dtype=torch.float64, device="cuda")
    ws = _workspace(lib, torch, n, kind)
    fn = getattr(lib, "sosat_" + fn_name)
    _lib.check(fn(ctypes.c_void_p(d_a.data_ptr()), ctypes.c_void_p(d_p.data_ptr()), n,
                  ctypes.c_void_p(field.data_ptr()), ctypes.c_void_p(phase.data_ptr()),
                  ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream(torch)))
    return _lib.to_numpy(torch.view_as_complex(field)), _lib.to_numpy(phase)


def b2a(beam, phase):
    """FFT angular space to aperture plane (opt_analyze.py:208-217).

    ``beam``: [N, N] (``abs`` is taken), ``phase`` in DEGREES.  Returns ``(aper_field complex128,
    aper_phase float64)`` with ``aper_field = fftshift(ifft2(|beam| exp(i phase pi/180)))`` --
    the reference computes ``fftshift(beam_complex)`` on line 213 and then discards it on line
    214, so the input is NOT shifted; kept.  Same precision as :func:`a2b` (fp32 information in a
    complex128 array, <= 1.5e-5 of peak)."""
    return _amp_phase_transform("b2a", beam, phase, _lib.DFT_B2A)


def _dft_kind(lib, torch, d_in, kind):
    """Centred transform of a complex64 CUDA tensor [N, N] through sosat_dft2_gemm_kind."""
    n = d_in.shape[0]
    out = torch.empty((n, n), dtype=torch.complex64, device="cuda")
    ws = _workspace(lib, torch, n, kind)
    _lib.check(lib.sosat_dft2_gemm_kind(ctypes.c_void_p(d_in.data_ptr()), n,
                                        ctypes.c_void_p(out.data_ptr()), int(kind),
                                        ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                        _stream(torch)))
    return out


def beam_convolve(x, y, beam, apert1, apert2, plots=0):
    """Convolve ``beam`` with the cos-tapered rectangular horn aperture (opt_analyze.py:60-88):
    ``ifftshift(ifft2(ifftshift(F(beam) * F(disc))))`` with ``F = fftshift . fft2 . fftshift``.
    ``x``, ``y``: [N, N] coordinate grids, ``beam``: complex [N, N].  Returns
    ``(x, y, beam_final complex128)``; ``plots`` is accepted and ignored."""
    lib = _lib.require_gpu()
    torch = _torch()
    xa = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    ya = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    bm = np.asarray(beam)
    if bm.ndim != 2 or bm.shape[0] != bm.shape[1] or xa.shape != bm.shape or ya.shape != bm.shape:
        raise ValueError("beam_convolve expects square [N, N] x, y and beam arrays of equal shape")
    n = bm.shape[0]
    # fp32 storage on the device: normalise so that small-valued beams keep their precision
    scale = float(np.max(np.abs(bm))) if bm.size else 0.0
    if not np.isfinite(scale) or scale == 0.0:
        scale = 1.0
    d_beam = torch.from_numpy(np.ascontiguousarray((bm / scale).astype(np.complex64))).to("cuda")
    d_x = torch.from_numpy(xa).to("cuda")
    d_y = torch.from_numpy(ya).to("cuda")
    disc = torch.empty((n, n), dtype=torch.complex64, device="cuda")
    _lib.check(lib.sosat_horn_aperture(ctypes.c_void_p(d_x.data_ptr()),
                                       ctypes.c_void_p(d_y.data_ptr()), n * n, float(apert1),
                                       float(apert2), ctypes.c_void_p(disc.data_ptr()),
                                       _stream(torch)))
    disc_fft = _dft_kind(lib, torch, disc, _lib.DFT_FORWARD)
    beam_fft = _dft_kind(lib, torch, d_beam, _lib.DFT_FORWARD)
    # the product can reach N^2 * N^2: rescale before the fp32 inverse pass
    peak = float(torch.view_as_real(disc_fft).abs().max())
    peak = peak if peak > 0 else 1.0
    conv = torch.empty_like(beam_fft)
    disc_fft /= peak
    _lib.check(lib.sosat_cmul(ctypes.c_void_p(beam_fft.data_ptr()),
                              ctypes.c_void_p(disc_fft.data_ptr()), n * n,
                              ctypes.c_void_p(conv.data_ptr()), _stream(torch)))
    final = _dft_kind(lib, torch, conv, _lib.DFT_INVERSE)
    return x, y, _lib.to_numpy(final).astype(np.complex128) * (scale * peak)


def far_field(b, dx_cm=None, lam_m=None, theta_x=None, theta_y=None, *, device: bool = False,
              out=None):
    """Far field of a complex near field ``b`` [N, N] (numpy complex or CUDA complex64 tensor).

    Without angles: the centred 2-D DFT ``B = fftshift(fft2(fftshift(b)))``; if the cell size
    ``dx_cm`` and the wavelength ``lam_m`` are given, returns ``(B, theta)`` with ``theta`` the
    true FFT-bin angles [rad] of both axes.

    With ``theta_x`` / ``theta_y`` (1-D arrays of M <= N angles [rad] each; needs ``dx_cm`` and
    ``lam_m``): the zoomed matrix DFT ``B[j, k] = sum_mn b[m, n] exp(-2 pi i (y_m theta_y[j] +
    x_n theta_x[k]) / lam)`` on exactly those angles, cell centres ``x_n = (n - ceil(N/2)) dx``
    (the convention under which the FFT-bin angles reproduce the plain transform).  Returns the
    ``[M, M]`` field.  ``device=True`` returns a CUDA complex64 tensor (no host copies); ``out``: a
    contiguous CUDA complex64 ``[N, N]`` tensor that receives the plain transform (e.g. a view of
    a buffer other ranks can read, ``dist.PeerPlanes.out``)."""
    lib = _lib.require_gpu()
    torch = _torch()
    if isinstance(b, np.ndarray):
        d_b = torch.from_numpy(np.ascontiguousarray(b.astype(np.complex64))).to("cuda")
    else:
        d_b = b.to(device="cuda", dtype=torch.complex64).contiguous()
    if d_b.dim() != 2 or d_b.shape[0] != d_b.shape[1]:
        raise ValueError("far_field expects a square [N, N] field")
    n = d_b.shape[0]
    if theta_x is not None or theta_y is not None:
        if theta_x is None or theta_y is None or dx_cm is None or lam_m is None:
            raise ValueError("zoomed far field needs theta_x, theta_y, dx_cm and lam_m")
        tx = np.ascontiguousarray(np.asarray(theta_x, dtype=np.float64).ravel())
        ty = np.ascontiguousarray(np.asarray(theta_y, dtype=np.float64).ravel())
        if tx.size != ty.size or not 1 <= tx.size <= n:
            raise ValueError("theta_x and theta_y must be 1-D arrays of the same length M, 1 <= M <= N")
        m = tx.size
        xs = np.ascontiguousarray((np.arange(n) - (n + 1) // 2) * (float(dx_cm) / 100.0))
        dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
        out = torch.empty((m, m), dtype=torch.complex64, device="cuda")
        ws = _workspace(lib, torch, n, 100)  # own workspace: its twiddles are rewritten per call
        _lib.check(lib.sosat_dft2_zoom(ctypes.c_void_p(d_b.data_ptr()), n,
                                       ctypes.c_void_p(out.data_ptr()), m, dp(xs), dp(xs), dp(ty),
                                       dp(tx), 1.0 / float(lam_m), ctypes.c_void_p(ws.data_ptr()),
                                       ws.numel(), _stream(torch)))
        torch.cuda.current_stream().synchronize()  # the host arrays above are read asynchronously
        return out if device else _lib.to_numpy(out)
    if out is None:
        out = torch.empty((n, n), dtype=torch.complex64, device="cuda")
    elif (not out.is_cuda or out.dtype != torch.complex64 or tuple(out.shape) != (n, n)
          or not out.is_contiguous()):
        raise ValueError("out must be a contiguous CUDA complex64 [N, N] tensor")
    ws = _workspace(lib, torch, n)
    _lib.check(lib.sosat_dft2_gemm(ctypes.c_void_p(d_b.data_ptr()), n,
                                   ctypes.c_void_p(out.data_ptr()),
                                   ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream(torch)))
    res = out if device else _lib.to_numpy(out)
    if dx_cm is not None and lam_m is not None:
        return res, far_field_angles(n, float(dx_cm) / 100.0, float(lam_m))
    return res


_batch_workspaces = {}


def far_field_batch(b, *, device: bool = False, pad: int = 0):
    """Centred 2-D DFT of a stack of near fields ``b`` [K, N, N] (numpy complex or CUDA complex64
    tensor), e.g. all frequencies or receivers of a sweep (``near_field_binned(...).b``): one
    transpose/split kernel and one launch per UMMA pass for the whole stack, so that sizes too
    small to fill the GPU on their own (N = 1024) run on the ``cta_group::2`` pair kernel.

    ``pad``: transform every field as if ``zero_pad`` had put ``pad`` zero cells on each side
    (finer angle sampling, output ``[K, N + 2 pad, N + 2 pad]``).  The zeros are neither stored
    nor multiplied (``sosat_dft2_gemm_batch_padded``: both passes contract only over the N columns
    of W that meet data -- 3/8 of the work at pad = N / 2).
    Returns complex64 (numpy, or a CUDA tensor with ``device=True``)."""
    lib = _lib.require_gpu()
    torch = _torch()
    if isinstance(b, np.ndarray):
        d_b = torch.from_numpy(np.ascontiguousarray(b.astype(np.complex64))).to("cuda")
    else:
        d_b = b.to(device="cuda", dtype=torch.complex64).contiguous()
    if d_b.dim() != 3 or d_b.shape[1] != d_b.shape[2]:
        raise ValueError("far_field_batch expects a [K, N, N] stack of square fields")
    pad = int(pad)
    if pad < 0:
        raise ValueError("pad must be >= 0")
    k, n_in = int(d_b.shape[0]), int(d_b.shape[1])
    n = n_in + 2 * pad
    out = torch.empty((k, n, n), dtype=torch.complex64, device="cuda")
    if k == 0:
        return out if device else _lib.to_numpy(out)
    if pad > 0 and ((n + 127) // 128 * 128) % 256 == 0:
        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream, "pad", n_in, pad, k)
        ws = _batch_workspaces.get(key)
        if ws is None:
            for old in [q for q in _batch_workspaces if q[:3] == key[:3]]:  # one padded stack per stream
                del _batch_workspaces[old]
            nbytes = int(lib.sosat_dft2_batch_padded_workspace_bytes(n_in, pad, k))
            ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            ptr = ws.data_ptr()
            weakref.finalize(ws, lambda p=ptr: lib.sosat_dft2_forget(ctypes.c_void_p(p)))
            _batch_workspaces[key] = ws
        _lib.check(lib.sosat_dft2_gemm_batch_padded(
            ctypes.c_void_p(d_b.data_ptr()), n_in, pad, k, ctypes.c_void_p(out.data_ptr()),
            _lib.DFT_FORWARD, ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream(torch)))
        return out if device else _lib.to_numpy(out)
    if pad > 0:  # sizes the compact form does not take: materialise the zeros
        d_b = torch.nn.functional.pad(d_b, (pad, pad, pad, pad))
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream, n, k)
    ws = _batch_workspaces.get(key)
    if ws is None:
        for old in [q for q in _batch_workspaces if q[:3] == key[:3]]:  # one stack size per n
            del _batch_workspaces[old]
        nbytes = int(lib.sosat_dft2_batch_workspace_bytes(n, k))
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        ptr = ws.data_ptr()
        weakref.finalize(ws, lambda p=ptr: lib.sosat_dft2_forget(ctypes.c_void_p(p)))
        _batch_workspaces[key] = ws
    _lib.check(lib.sosat_dft2_gemm_batch(ctypes.c_void_p(d_b.data_ptr()), n, k,
                                         ctypes.c_void_p(out.data_ptr()), _lib.DFT_FORWARD,
                                         ctypes.c_void_p(ws.data_ptr()), ws.numel(), _stream(torch)))
    return out if device else _lib.to_numpy(out)


def far_field_angles(n: int, dx_m: float, lam_m: float):
    """True FFT-bin angles of the centred transform: theta_k = (k - floor(N/2)) lam / (N dx)
    (from opt_analyze.py:283-294; the labels of ``coords_spat_to_ang`` are these stretched by
    N/(N-1))."""
    return (np.arange(n) - n // 2) * lam_m / (n * dx_m)


def far_field_irregular(x, y, b, theta_x, theta_y, lam):
    """Direct-sum far field of irregular samples: ``B_j = sum_p b_p exp(-2 pi i (x_p thx_j +
    y_p thy_j) / lam)``; ``x, y, lam`` in the same length unit, angles in radians.  The sign
    follows the reference's code path (numpy FFT), not README.md:24.

    ``x`` may instead be a CUDA float32 tensor ``[P, 4]`` of samples ``{x, y, Re b, Im b}`` as
    returned by ``ray_trace.near_field_samples(..., device=True)`` (then ``y`` and ``b`` must be
    ``None``): the samples stay on the device."""
    lib = _lib.require_gpu()
    torch = _torch()
    tx = np.asarray(theta_x, dtype=np.float64)
    ty = np.asarray(theta_y, dtype=np.float64)
    if tx.shape != ty.shape:
        raise ValueError("theta_x and theta_y must have the same shape")
    if not isinstance(x, np.ndarray) and hasattr(x, "is_cuda"):
        if y is not None or b is not None:
            raise ValueError("with a device sample tensor, y and b must be None")
        d_pts = x.to(device="cuda", dtype=torch.float32).contiguous()
        if d_pts.dim() != 2 or d_pts.shape[1] != 4:
            raise ValueError("device samples must be [P, 4] float32 {x, y, Re b, Im b}")
        n_pts = d_pts.shape[0]
    else:
        x = np.asarray(x, dtype=np.float64).ravel()
        y = np.asarray(y, dtype=np.float64).ravel()
        b = np.asarray(b, dtype=np.complex128).ravel()
        if not (x.size == y.size == b.size):
            raise ValueError("x, y, b must have the same number of samples")
        n_pts = x.size
        pts = np.stack([x, y, b.real, b.imag], axis=1).astype(np.float32)
        d_pts = torch.from_numpy(np.ascontiguousarray(pts)).to("cuda") if n_pts else None
    if n_pts == 0 or tx.size == 0:
        return np.zeros(tx.shape, dtype=np.complex128)
    th = np.stack([tx.ravel(), ty.ravel()], axis=1).astype(np.float32)
    d_th = torch.from_numpy(np.ascontiguousarray(th)).to("cuda")
    out = torch.empty((th.shape[0], 2), dtype=torch.float32, device="cuda")
    _lib.check(lib.sosat_nudft_direct(ctypes.c_void_p(d_pts.data_ptr()), n_pts,
                                      ctypes.c_void_p(d_th.data_ptr()), th.shape[0],
                                      1.0 / float(lam), ctypes.c_void_p(out.data_ptr()),
                                      _stream(torch)))
    res = _lib.to_numpy(out)
    return (res[:, 0] + 1j * res[:, 1]).reshape(tx.shape)
