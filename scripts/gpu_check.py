"""Developer bring-up script (run on the GPU box): parity + timing diagnostics, verbose."""
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage_trace():
    import torch
    from conftest import golden, tele_geo_from_golden, phase_diff
    from sosat_optics_b200 import _lib, ray_trace, geometry, ot_geo
    lib = _lib.require_gpu()
    name = ctypes.create_string_buffer(256)
    sm, nsm, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    _lib.check(lib.sosat_device_info(name, ctypes.byref(sm), ctypes.byref(nsm), ctypes.byref(mem)))
    print("device", name.value.decode(), "sm", sm.value, "SMs", nsm.value, "mem GB", mem.value / 2**30)
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.sosat_measure_fp64_peak(ctypes.byref(fl), ctypes.byref(ms)))
    print("fp64 FMA peak: %.2f TFLOP/s (%.3f ms)" % (fl.value / 1e12, ms.value))

    g = golden("trace_cfg1_n60_106ghz.npz")
    t = tele_geo_from_golden(g)
    ref = g["out"]
    for var in (0, 2, 6):
        os.environ["SOSAT_TRACE_VARIANT"] = str(var)
        out = ray_trace.rx_to_lyot([0, 0, 0], t, False, "b")
        vr, vo = ref[4] != 0, out[4] != 0
        m = vr & vo
        d = [np.abs(out[r][m] - ref[r][m]).max() for r in range(11)]
        print("variant", var, "mask equal", bool((vr == vo).all()), "n_valid", int(vo.sum()),
              "max|d| pos %.2e %.2e opl %.2e amp(rel) %.2e nrm %.2e dir %.2e" % (
                  d[0], d[2], d[3], np.abs(out[4][m] / ref[4][m] - 1).max(), max(d[5:8]), max(d[8:11])))
    os.environ["SOSAT_TRACE_VARIANT"] = "0"
    sb = ray_trace.getNearField(t, [0, 0, 0])
    print("getNearField: n", len(sb.a_sim), "ref n", len(g["sb_a_sim"]),
          "max dx %.2e dy %.2e da %.2e dp %.2e tan %.2e" % (
              np.abs(sb.x_sim - g["sb_x_sim"]).max(), np.abs(sb.y_sim - g["sb_y_sim"]).max(),
              np.abs(sb.a_sim - g["sb_a_sim"]).max(), phase_diff(sb.p_sim, g["sb_p_sim"]).max(),
              np.abs(sb.tan_og - g["sb_tan_og"]).max()))
    print("fwhm", ray_trace.get_fwhm(sb), "ref", g["fwhm"])

    g4 = golden("trace_cfg4_offaxis_n32_90ghz.npz")
    t4 = tele_geo_from_golden(g4)
    for i, rx in enumerate(g4["rx"]):
        out = ray_trace.rx_to_lyot(list(rx), t4, False, "b")
        ref = g4["out"][i]
        vr, vo = ref[4] != 0, out[4] != 0
        m = vr & vo
        print("offaxis", rx, "mask equal", bool((vr == vo).all()), int(vr.sum()),
              "opl %.2e pos %.2e" % (np.abs(out[3][m] - ref[3][m]).max(), np.abs(out[0][m] - ref[0][m]).max()))

    # binned vs rows-derived binning, small
    from oracle import oracle
    t.n_scan = 200
    out = ray_trace.rx_to_lyot([0, 0, 0], t)
    kk = float(np.asarray(t.k).ravel()[0])
    re, im, cnt, sums = oracle.bin_near_field(out, kk, 64, 420.0, 0.0)
    bf = ray_trace.near_field_binned(t, [0, 0, 0], grid_n=64, extent_cm=42.0, keep_sums=True)
    print("bin: cnt equal", bool((bf.count[0] == cnt).all()), "max|dre| %.2e max|dim| %.2e (max |re| %.2e)" % (
        np.abs(bf.re[0, 0] - re).max(), np.abs(bf.im[0, 0] - im).max(), np.abs(re).max()),
        "stats", bf.stats[0], "oracle", sums)

    # slow-path fraction for off-axis feeds, default schedule
    t.n_scan = 2000
    for rx in ([0, 0, 0], [0, 0, 150], [150, 0, 0], [-120, 0, 80], [106, 0, 106], [0, 0, 170]):
        st = ray_trace.trace_stats(t, rx)[0]
        print("rx", rx, "n_valid %d n_slow %d (%.4f%%) bracket_fail %d" % (st[0], st[5], 100 * st[5] / max(st[0], 1), st[6]))
    # timing: Newton schedule x resident CTAs/SM
    t.n_scan = 8192
    spec = [(t.k, t.th_fwhp_x, t.th_fwhp_y)]
    n = t.n_scan ** 2
    for minb in (3, 4):
        os.environ["SOSAT_TRACE_MINB"] = str(minb)
        for var in (0, 2, 6):
            os.environ["SOSAT_TRACE_VARIANT"] = str(var)
            planes = None
            best = 1e9
            for rep in range(4):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                re_, im_, cnt_, st_ = ray_trace.trace_bin_device(t, [[0, 0, 0]], spec, 2048, 420.0, planes=planes)
                e1.record()
                torch.cuda.synchronize()
                planes = (re_, im_, cnt_)
                if rep:
                    best = min(best, e0.elapsed_time(e1))
            print("minb", minb, "variant", var, "bin-mode n_scan=8192: %.2f ms -> %.3e rays/s" % (best, n / best * 1e3),
                  "stats", st_.cpu().numpy()[0][[0, 5, 6, 7]], flush=True)
    os.environ["SOSAT_TRACE_VARIANT"] = "0"
    os.environ.pop("SOSAT_TRACE_MINB", None)


def stage_dft():
    import torch
    from conftest import golden
    from sosat_optics_b200 import opt_analyze
    rng = np.random.default_rng(1)
    for n in (128, 256, 240, 65, 1000, 1024):
        b = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
        ref = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(b.astype(np.complex128))))
        out = opt_analyze.far_field(b)
        err = np.abs(out - ref).max() / np.abs(ref).max()
        print("dft2 n=%d rel-to-peak err %.3e" % (n, err), flush=True)
    g = golden("a2b_small.npz")
    for n in (16, 64, 65, 100, 129):
        bc, ph = opt_analyze.a2b(g[f"amp_{n}"], g[f"phi_{n}"])
        print("a2b n=%d err %.3e" % (n, np.abs(bc - g[f"beam_{n}"]).max() / np.abs(g[f"beam_{n}"]).max()))
    d = torch.randn(1024, 1024, dtype=torch.complex64, device="cuda")
    for n in (1024, 2048):
        d = torch.randn(n, n, dtype=torch.complex64, device="cuda")
        opt_analyze.far_field(d, device=True)
        torch.cuda.synchronize()
        ts = []
        for rep in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt_analyze.far_field(d, device=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("far_field n=%d: best %.3f ms median %.3f ms" % (n, min(ts), sorted(ts)[len(ts) // 2]))


def stage_nudft():
    from sosat_optics_b200 import opt_analyze
    from oracle import oracle
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-0.2, 0.2, 5000), rng.uniform(-0.2, 0.2, 5000)
    b = rng.standard_normal(5000) + 1j * rng.standard_normal(5000)
    th = np.linspace(-0.02, 0.02, 33)
    TX, TY = np.meshgrid(th, th)
    out = opt_analyze.far_field_irregular(x, y, b, TX, TY, 0.00333)
    ref = oracle.nudft_direct(x, y, b, TX.ravel(), TY.ravel(), 0.00333).reshape(TX.shape)
    print("nudft err rel peak %.3e" % (np.abs(out - ref).max() / np.abs(ref).max()))


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage == "all":
        for s in ("trace", "dft", "nudft"):
            print("=====", s, flush=True)
            r = subprocess.run(["timeout", "300", sys.executable, __file__, s])
            print("exit", r.returncode, flush=True)
    else:
        {"trace": stage_trace, "dft": stage_dft, "nudft": stage_nudft}[stage]()
