"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REAL reference.

Run in the development container (needs /root/reference):

    python oracle/make_golden.py            # everything (~8 min, single core)
    python oracle/make_golden.py trace a2b  # a subset

The reference is imported unmodified behind the shims in oracle/ref_shim.py.  Every
array stored here is an output of the reference's own functions
(ray_trace.rx_to_lyot / getNearField / get_fwhm / sim2d / sim_data / ff_fwhm_est,
opt_analyze.zero_pad / a2b / coords_spat_to_ang) on the BASELINE.json configurations.
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_feedhorn_table, load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
warnings.simplefilter("ignore")


def _geo(ot_geo, opt_analyze, fw, *, n_scan, y_off, ghz=None, lam=None, sqrt2=False,
         as_array=True):
    g = ot_geo.SatGeo()
    g.n_scan = n_scan
    g.y_source = ot_geo.y_lyot + y_off
    g.lambda_ = opt_analyze.ghz_to_m(ghz) if lam is None else lam
    g.k = 2 * np.pi / g.lambda_
    row = int(opt_analyze.m_to_ghz(g.lambda_)) if ghz is None else int(ghz)
    fx = np.deg2rad(fw[np.where(fw[:, 0] == row), 1])
    fy = np.deg2rad(fw[np.where(fw[:, 0] == row), 2])
    if sqrt2:
        fx, fy = fx * np.sqrt(2), fy * np.sqrt(2)
    if not as_array:
        fx, fy = fx[0][0], fy[0][0]
    g.th_fwhp_x, g.th_fwhp_y = fx, fy
    return g, row


def _sb_dict(sb, prefix="sb_"):
    return {prefix + k: getattr(sb, k) for k in ("x_sim", "y_sim", "a_sim", "p_sim", "tan_og")}


def gen_feedhorn(mods, fw):
    np.savez_compressed(os.path.join(GOLD, "feedhorn_fwhm.npz"), table=fw)


def gen_trace(mods, fw):
    ot_geo, ray_trace, opt_analyze = mods
    # cfg 1: sat_nearfield.ipynb cell 1 (n_scan=60, 106 GHz, y_source=y_lyot+300, boresight)
    g, row = _geo(ot_geo, opt_analyze, fw, n_scan=60, y_off=300, ghz=106)
    t = time.time()
    out = ray_trace.rx_to_lyot([0, 0, 0], g, False, "b")
    dt = time.time() - t
    sb = ray_trace.getNearField(g, [0, 0, 0], plot=False)
    fwhm = ray_trace.get_fwhm(sb)
    np.savez_compressed(
        os.path.join(GOLD, "trace_cfg1_n60_106ghz.npz"),
        out=out[:11], rx=np.zeros(3), n_scan=60, y_source=g.y_source, lambda_=g.lambda_,
        k=g.k, th_fwhp_x=g.th_fwhp_x, th_fwhp_y=g.th_fwhp_y, table_row=row,
        fwhm=np.array(fwhm), beam_cent=np.array(ray_trace.get_beam_cent(sb)),
        angle_out=np.array(ray_trace.get_angle_out(sb)), ref_seconds=dt, **_sb_dict(sb))
    print("cfg1", dt, "s  fwhm", fwhm, "valid", (out[4] != 0).sum())

    # cfg 1 at 90 GHz (BASELINE.json wording), smaller bundle
    g, row = _geo(ot_geo, opt_analyze, fw, n_scan=40, y_off=300, ghz=90)
    out = ray_trace.rx_to_lyot([0, 0, 0], g, False, "b")
    sb = ray_trace.getNearField(g, [0, 0, 0], plot=False)
    np.savez_compressed(
        os.path.join(GOLD, "trace_cfg1_n40_90ghz.npz"),
        out=out[:11], rx=np.zeros(3), n_scan=40, y_source=g.y_source, lambda_=g.lambda_,
        k=g.k, th_fwhp_x=g.th_fwhp_x, th_fwhp_y=g.th_fwhp_y, table_row=row, **_sb_dict(sb))

    # cfg 4: off-axis receivers (focal-plane sweep), n_scan=32, 90 GHz, y_source=y_lyot+200
    rxs = np.array([[100, 0, 60], [-120, 0, 80], [0, 0, 150], [150, 0, 0], [60, 0, -90],
                    [-30, 0, -140], [106.07, 0, 106.07]], dtype=float)
    g, row = _geo(ot_geo, opt_analyze, fw, n_scan=32, y_off=200, ghz=90, as_array=False)
    outs, sbs = [], {}
    for i, rx in enumerate(rxs):
        outs.append(ray_trace.rx_to_lyot(list(rx), g, False, "b")[:11])
        sb = ray_trace.getNearField(g, list(rx), plot=False)
        sbs.update(_sb_dict(sb, prefix=f"sb{i}_"))
        print("offaxis", rx, "valid", (outs[-1][4] != 0).sum())
    np.savez_compressed(
        os.path.join(GOLD, "trace_cfg4_offaxis_n32_90ghz.npz"),
        out=np.stack(outs), rx=rxs, n_scan=32, y_source=g.y_source, lambda_=g.lambda_,
        k=g.k, th_fwhp_x=g.th_fwhp_x, th_fwhp_y=g.th_fwhp_y, table_row=row, **sbs)


def gen_reqs(mods, fw):
    """cfg 3: sat_reqs.ipynb cell 1 -- near-field FWHM vs frequency, n_scan=150."""
    ot_geo, ray_trace, opt_analyze = mods
    freqs = np.arange(70, 180, 10)
    res = np.zeros((len(freqs), 2))
    nvalid = np.zeros(len(freqs), dtype=np.int64)
    cents = np.zeros((len(freqs), 2))
    fxy = np.zeros((len(freqs), 2))
    for i, f in enumerate(freqs):
        g, row = _geo(ot_geo, opt_analyze, fw, n_scan=150, y_off=200, ghz=f, as_array=False)
        t = time.time()
        sb = ray_trace.getNearField(g, [0, 0, 0], plot=False)
        res[i] = ray_trace.get_fwhm(sb)
        nvalid[i] = len(sb.a_sim)
        cents[i] = ray_trace.get_beam_cent(sb)
        fxy[i] = (g.th_fwhp_x, g.th_fwhp_y)
        print("reqs", f, res[i], nvalid[i], "%.1fs" % (time.time() - t), flush=True)
    np.savez_compressed(os.path.join(GOLD, "reqs_cfg3_n150.npz"), freqs=freqs, fwhm=res,
                        n_valid=nvalid, beam_cent=cents, th_fwhp=fxy,
                        y_source=ot_geo.y_lyot + 200, n_scan=150)


def _farfield_chain(mods, g, stage_range, step, pix_num=100):
    ot_geo, ray_trace, opt_analyze = mods
    sb = ray_trace.getNearField(g, [0, 0, 0], plot=False)
    sb2 = ray_trace.sim2d(sb)
    d = ray_trace.sim_data(sb2, stage_range, step)
    x_new, y_new, beam_data = opt_analyze.zero_pad(d.x_data, d.y_data, d.a_data, pix_num)
    x_data, y_data, phase_data = opt_analyze.zero_pad(d.x_data, d.y_data, d.p_data, pix_num)
    beam_fft, phase_fft = opt_analyze.a2b(beam_data, np.rad2deg(phase_data))
    x_ang, y_ang = opt_analyze.coords_spat_to_ang(
        x_data / 1e2, y_data / 1e2, opt_analyze.m_to_ghz(g.lambda_))
    fwhm = ray_trace.ff_fwhm_est(beam_fft, x_ang, y_ang)
    return dict(sb=sb, sb2=sb2, d=d, beam_data=beam_data, phase_data=phase_data,
                x_data=x_data, y_data=y_data, beam_fft=beam_fft, phase_fft=phase_fft,
                x_ang=x_ang, y_ang=y_ang, fwhm=fwhm)


def gen_farfield(mods, fw):
    ot_geo, ray_trace, opt_analyze = mods
    # cfg 2: sat_farfield.ipynb cells 1-6 (n_scan=100, lambda=2.3 mm -> row 130, 240^2)
    g, row = _geo(ot_geo, opt_analyze, fw, n_scan=100, y_off=300, lam=0.0023)
    c = _farfield_chain(mods, g, 80, 2)
    print("cfg2 far-field FWHM", c["fwhm"], "arcmin; shape", c["beam_fft"].shape)
    np.savez_compressed(
        os.path.join(GOLD, "farfield_cfg2_240.npz"),
        a_data=c["d"].a_data, p_data=c["d"].p_data, x_data=c["d"].x_data,
        y_data=c["d"].y_data, pix_num=100, x_pad=c["x_data"], y_pad=c["y_data"],
        beam_pad=c["beam_data"], phase_pad=c["phase_data"], beam_fft=c["beam_fft"],
        phase_fft=c["phase_fft"], x_ang_row=c["x_ang"][0], y_ang_col=c["y_ang"][:, 0],
        fwhm=c["fwhm"], lambda_=g.lambda_, table_row=row,
        sb2_a=c["sb2"].a_sim, sb2_p=c["sb2"].p_sim, sb2_x=c["sb2"].x_sim,
        sb2_y=c["sb2"].y_sim)

    # cfg 2 loop: sat_farfield.ipynb cell 8 at 90 GHz (n_scan=150, 800^2 padded to 1000^2).
    # The 1000^2 arrays are too large to commit; store the 100x100 sim2d grid the
    # reference produced, the known answer, and samples of the reference's a2b output.
    g, row = _geo(ot_geo, opt_analyze, fw, n_scan=150, y_off=300, ghz=90, sqrt2=True)
    c = _farfield_chain(mods, g, 200, .25)
    bf = c["beam_fft"]
    n = bf.shape[0]
    pk = np.unravel_index(np.argmax(abs(bf)), bf.shape)
    rng = np.random.default_rng(5)
    samp = rng.integers(0, n, size=(4096, 2))
    print("cfg2-loop 90 GHz far-field FWHM", c["fwhm"], "arcmin; shape", bf.shape)
    np.savez_compressed(
        os.path.join(GOLD, "farfield_cfg2loop_1000_90ghz.npz"),
        sb2_a=c["sb2"].a_sim, sb2_p=c["sb2"].p_sim, sb2_x=c["sb2"].x_sim,
        sb2_y=c["sb2"].y_sim, stage_range=200, step=.25, pix_num=100, fwhm=c["fwhm"],
        lambda_=g.lambda_, table_row=row, n=n, peak_index=np.array(pk), peak_value=bf[pk],
        center_block=bf[n // 2 - 24:n // 2 + 24, n // 2 - 24:n // 2 + 24],
        sample_index=samp, sample_value=bf[samp[:, 0], samp[:, 1]],
        sample_phase=c["phase_fft"][samp[:, 0], samp[:, 1]],
        a_data_checksum=float(np.sum(c["d"].a_data)), p_data_checksum=float(np.sum(c["d"].p_data)),
        a_data_diag=np.diag(c["d"].a_data), p_data_diag=np.diag(c["d"].p_data))


def gen_a2b(mods, fw):
    """a2b / zero_pad / coords_spat_to_ang on small synthetic fields, incl. odd N."""
    ot_geo, ray_trace, opt_analyze = mods
    rng = np.random.default_rng(2)
    store = {}
    for n in (16, 64, 65, 100, 129):
        amp = rng.uniform(0, 1, (n, n))
        amp[rng.uniform(size=(n, n)) < 0.1] *= -1  # a2b takes abs()
        phi = rng.uniform(-180, 180, (n, n))
        bc, ph = opt_analyze.a2b(amp, phi)
        store[f"amp_{n}"], store[f"phi_{n}"] = amp, phi
        store[f"beam_{n}"], store[f"phase_{n}"] = bc, ph
    # smooth aperture with zero_pad + coordinates
    x = np.arange(-10, 10, 0.5)
    X, Y = np.meshgrid(x, x)
    R = np.hypot(X, Y)
    amp = np.exp(-R ** 2 / (2 * 4.0 ** 2)) * (R < 9)
    amp[3, 5] = np.nan  # zero_pad maps NaN -> 0 (opt_analyze.py:31)
    ph = 0.3 * np.cos(X / 3) + 0.05 * (R / 9) ** 4
    xp, yp, ap = opt_analyze.zero_pad(X, Y, amp, 12)
    _, _, pp = opt_analyze.zero_pad(X, Y, ph, 12)
    bc, phs = opt_analyze.a2b(ap, np.rad2deg(pp))
    xa, ya = opt_analyze.coords_spat_to_ang(xp / 1e2, yp / 1e2, 90.0)
    store.update(zp_x=X, zp_y=Y, zp_amp=amp, zp_ph=ph, zp_pts=12, zp_xo=xp, zp_yo=yp,
                 zp_ao=ap, zp_po=pp, zp_beam=bc, zp_phase=phs, zp_xang=xa, zp_yang=ya,
                 zp_fwhm=ray_trace.ff_fwhm_est(bc, xa, ya))
    np.savez_compressed(os.path.join(GOLD, "a2b_small.npz"), **store)


def gen_convolve(mods, fw):
    """b2a, beam_convolve and coords_ang_to_spat (rows f3/f4) on small fields, odd and even N."""
    ot_geo, ray_trace, opt_analyze = mods
    rng = np.random.default_rng(7)
    store = {}
    for n in (16, 65, 96, 129):
        amp = rng.uniform(0, 1, (n, n))
        amp[rng.uniform(size=(n, n)) < 0.1] *= -1  # b2a takes abs()
        phi = rng.uniform(-180, 180, (n, n))
        af, ap = opt_analyze.b2a(amp, phi)
        store[f"b2a_amp_{n}"], store[f"b2a_phi_{n}"] = amp, phi
        store[f"b2a_field_{n}"], store[f"b2a_phase_{n}"] = af, ap
    # a2b -> b2a on a smooth aperture (what a holography analysis does)
    for n, dx in ((64, 0.5), (101, 0.4)):
        x = (np.arange(n) - n // 2) * dx
        X, Y = np.meshgrid(x, x)
        R = np.hypot(X, Y)
        beam = np.exp(-R ** 2 / (2 * 5.0 ** 2)) * np.exp(1j * (0.4 * np.cos(X / 4) + 0.02 * Y))
        beam[R > 14] = 0
        for (a1, a2) in ((2.0, 3.0), (3.5, 1.5), (2.0, 2.0)):
            _, _, bf = opt_analyze.beam_convolve(X, Y, beam, a1, a2, 0)
            store[f"conv_{n}_{a1}_{a2}"] = bf
        store[f"conv_x_{n}"], store[f"conv_y_{n}"], store[f"conv_beam_{n}"] = X, Y, beam
    th = np.linspace(-0.2, 0.2, 81)
    TX, TY = np.meshgrid(th, 1.5 * th)
    xs, ys = opt_analyze.coords_ang_to_spat(TX, TY, 90.0)
    store.update(a2s_tx=TX, a2s_ty=TY, a2s_x=xs, a2s_y=ys)
    np.savez_compressed(os.path.join(GOLD, "b2a_convolve_small.npz"), **store)


def gen_ot_geo(mods, fw):
    """The reference's Python-level surface functions (sag, slope, frame maps) at seeded points,
    incl. radii beyond the conic domain of lens 1b (NaN) -- ot_geo.py:49-458."""
    ot_geo, ray_trace, opt_analyze = mods
    rng = np.random.default_rng(11)
    x = rng.uniform(-320, 320, 257)
    y = rng.uniform(-320, 320, 257)
    z = rng.uniform(-50, 900, 257)
    store = dict(x=x, y=y, z=z)
    with np.errstate(invalid="ignore"):
        for name in ("1a", "1b", "2a", "2b", "3a", "3b"):
            store[f"z{name}"] = getattr(ot_geo, f"z{name}")(x, y)
            gx, gy = getattr(ot_geo, f"d_z{name}")(x, y)
            store[f"d_z{name}_x"], store[f"d_z{name}_y"] = gx, gy
            store[f"m{name}_into_tele"] = np.array(getattr(ot_geo, f"m{name}_into_tele")(x, y, z))
            # the reference shifts y in place: hand it copies
            store[f"tele_into_m{name}"] = np.array(
                getattr(ot_geo, f"tele_into_m{name}")(x.copy(), y.copy(), z.copy()))
    np.savez_compressed(os.path.join(GOLD, "ot_geo_surfaces.npz"), **store)


GENS = dict(ot_geo=gen_ot_geo, convolve=gen_convolve, feedhorn=gen_feedhorn, trace=gen_trace, a2b=gen_a2b, farfield=gen_farfield,
            reqs=gen_reqs)

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    mods = load_reference()
    fw = load_feedhorn_table()
    which = sys.argv[1:] or list(GENS)
    for w in which:
        t = time.time()
        GENS[w](mods, fw)
        print(f"[{w}] done in {time.time() - t:.1f}s", flush=True)
