"""CPU: the C-ABI library loads without a GPU, exports every symbol include/sosat_b200.h
declares, its structs have the documented layout, and the product path refuses to compute on
the CPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from sosat_optics_b200 import _lib, geometry, opt_analyze, ot_geo, ray_trace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sosat_b200.h")
no_gpu = not torch.cuda.is_available()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sosat_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sosat_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    assert lib.sosat_abi_version() == 1
    assert ctypes.sizeof(_lib.Surface) == 10 * 8
    assert ctypes.sizeof(_lib.Geom) == 6 * 80 + 5 * 8
    assert ctypes.sizeof(_lib.Freq) == 3 * 8
    hdr = open(HEADER).read()
    assert "#define SOSAT_OUT_ROWS 17" in hdr and _lib.OUT_ROWS == 17
    assert "#define SOSAT_MAX_FREQS 64" in hdr and _lib.MAX_FREQS == 64
    assert "SOSAT_NUM_STATS = 8" in hdr and _lib.NUM_STATS == 8


def test_workspace_size_is_pure_host_arithmetic():
    lib = _lib.load()
    assert lib.sosat_dft2_workspace_bytes(1024) == 256 + 56 * 1024 * 1024
    assert lib.sosat_dft2_workspace_bytes(1000) == lib.sosat_dft2_workspace_bytes(1024)
    assert lib.sosat_dft2_workspace_bytes(241) == 256 + 56 * 256 * 256
    assert lib.sosat_dft2_workspace_bytes(0) == 0


@pytest.mark.skipif(not no_gpu, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.sosat_device_count() == 0
    assert lib.sosat_set_device(0) == -5  # SOSAT_ERR_NO_DEVICE
    assert b"no CPU path" in lib.sosat_last_error()
    t = ot_geo.SatGeo()
    with pytest.raises(_lib.SosatError):
        ray_trace.rx_to_lyot([0, 0, 0], t, False, "b")
    with pytest.raises(_lib.SosatError):
        ray_trace.getNearField(t, [0, 0, 0])
    with pytest.raises(_lib.SosatError):
        opt_analyze.a2b(np.ones((8, 8)), np.zeros((8, 8)))
    with pytest.raises(_lib.SosatError):
        ray_trace.near_field_binned(t, [0, 0, 0], grid_n=8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sosat_optics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "build.py" \
                    or "import oracle" not in text and "from oracle" not in text
                assert "import oracle" not in text and "from oracle" not in text


def test_product_does_not_import_scipy():
    """Round 1's sim2d / sim_data were host scipy; since round 2 they run on the device and the
    product package must not import scipy anywhere (tests and the oracle may)."""
    import re
    pkg = os.path.join(ROOT, "sosat_optics_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            text = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(import|from)\s+scipy", text, flags=re.M), f
    # and the no-GPU error comes before any work in the two glue functions
    from types import SimpleNamespace
    sb = SimpleNamespace(x_sim=np.zeros(4), y_sim=np.zeros(4), a_sim=np.zeros(4), p_sim=np.zeros(4),
                         ray_index=np.arange(4), n_scan=2)
    with pytest.raises(_lib.SosatError):
        ray_trace.sim2d(sb)
    grid = SimpleNamespace(a_sim=np.zeros((8, 8)), p_sim=np.zeros((8, 8)),
                           x_sim=np.mgrid[0:1:8j, 0:1:8j][0], y_sim=np.mgrid[0:1:8j, 0:1:8j][1])
    with pytest.raises(_lib.SosatError):
        ray_trace.sim_data(grid, 2, 0.5)


def test_geometry_marshalling():
    t = ot_geo.SatGeo()
    t.y_source = ot_geo.y_lyot + 300
    t.th_fwhp_x = np.array([[0.564254]])  # notebooks assign (1,1) arrays
    t.th_fwhp_y = np.float64(0.6)
    g = geometry.make_geom(t)
    assert g.y_lyot == ot_geo.y_lyot and g.y_source == ot_geo.y_lyot + 300
    assert [g.surf[i].n_in for i in range(6)] == [1.0, 3.416, 1.0, 3.416, 1.0, 3.416]
    assert [g.surf[i].n_out for i in range(6)] == [3.416, 1.0, 3.416, 1.0, 3.416, 1.0]
    assert g.surf[4].k > -1 and all(g.surf[i].k < -1 for i in (0, 1, 2, 3, 5))
    assert (g.clip_l3, g.clip_stop, g.cone) == (196.0, 210.0, 0.4)
    f = geometry.make_freq(t)
    assert f.fwhp_x == 0.564254 and f.fwhp_y == 0.6
    assert f.k_over_2e3 == t.k / 1e3 / 2
    with pytest.raises(ValueError):
        geometry.scalar(np.zeros(2))
    # a user may change the refractive index on the instance
    t.n_si = 3.38
    assert geometry.make_geom(t).surf[0].n_out == 3.38


def test_peer_reduce_argument_checks_need_no_gpu():
    """sosat_bins_reduce_field validates its arguments before touching CUDA: the error codes and
    messages of the multi-GPU entry point can be checked on a CPU-only machine."""
    import ctypes
    lib = _lib.load()
    one = (ctypes.c_void_p * 1)(ctypes.c_void_p(256))
    many = (ctypes.c_void_p * 17)(*[ctypes.c_void_p(256)] * 17)

    def call(peers, n_peers, begin=0, end=8, re_off=0, outs=one, n_out=1):
        return lib.sosat_bins_reduce_field(peers, n_peers, re_off, 256, 512, 768, begin, end, 1.0, 0.0,
                                           outs, n_out, 1024, None, None)

    assert call(None, 1) != 0 and b"null pointer" in lib.sosat_last_error()
    assert call(one, 0) != 0 and b"between 1 and 16" in lib.sosat_last_error()
    assert call(many, 17) != 0 and b"between 1 and 16" in lib.sosat_last_error()
    assert call(one, 1, outs=many, n_out=17) != 0 and b"between 1 and 16" in lib.sosat_last_error()
    assert call(one, 1, begin=8, end=4) != 0 and b"bad cell range" in lib.sosat_last_error()
    assert call(one, 1, begin=2, end=8) != 0 and b"multiple of 4" in lib.sosat_last_error()
    assert call(one, 1, re_off=8) != 0 and b"misaligned plane offset" in lib.sosat_last_error()
    null_base = (ctypes.c_void_p * 1)(None)
    assert call(null_base, 1) != 0 and b"peer base null or misaligned" in lib.sosat_last_error()
    handle = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
    assert lib.sosat_peer_export(None, handle) != 0 and b"null pointer" in lib.sosat_last_error()
    assert lib.sosat_peer_open(None, None) != 0 and b"null pointer" in lib.sosat_last_error()
    assert lib.sosat_peer_free(None) == 0 and lib.sosat_peer_close(None) == 0
