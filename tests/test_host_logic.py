"""CPU: host-side helpers of the drop-in layer against goldens produced by the reference, and
the multi-GPU sharding / reduce logic on gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import golden

from sosat_optics_b200 import dist as sdist
from sosat_optics_b200 import geometry, opt_analyze, ray_trace


def test_zero_pad_and_coords_match_reference():
    g = golden("a2b_small.npz")
    xo, yo, ao = opt_analyze.zero_pad(g["zp_x"], g["zp_y"], g["zp_amp"], int(g["zp_pts"]))
    assert np.array_equal(xo, g["zp_xo"]) and np.array_equal(yo, g["zp_yo"])
    assert np.array_equal(ao, g["zp_ao"])
    xa, ya = opt_analyze.coords_spat_to_ang(xo / 1e2, yo / 1e2, 90.0)
    assert np.array_equal(xa, g["zp_xang"]) and np.array_equal(ya, g["zp_yang"])
    # label spacing is the true FFT-bin spacing stretched by N/(N-1)
    n = xa.shape[0]
    dx = abs(xo[0, 0] - xo[0, 1]) / 1e2
    true = opt_analyze.far_field_angles(n, dx, opt_analyze.ghz_to_m(90.0))
    assert np.allclose(np.diff(xa[0]), np.diff(true) * n / (n - 1), rtol=1e-9)
    assert true[n // 2] == 0.0


def test_unit_helpers():
    assert opt_analyze.ghz_to_m(150) == 0.002 and opt_analyze.m_to_ghz(0.002) == 150
    assert int(opt_analyze.m_to_ghz(0.0023)) == 130  # the sat_farfield notebook's table row
    assert opt_analyze.rad_to_arcmin(np.pi / 180) == pytest.approx(60.0)


def test_fwhm_helpers_on_reference_arrays():
    g = golden("trace_cfg1_n60_106ghz.npz")
    sb = ray_trace.SimOutput(g["sb_x_sim"], g["sb_y_sim"], g["sb_a_sim"], g["sb_p_sim"],
                             g["sb_tan_og"])
    assert np.allclose(ray_trace.get_fwhm(sb), g["fwhm"], atol=1e-12)
    assert np.allclose(ray_trace.get_beam_cent(sb), g["beam_cent"], atol=1e-12)
    assert np.allclose(ray_trace.get_angle_out(sb), g["angle_out"], atol=1e-12)
    f = golden("farfield_cfg2_240.npz")
    xa, ya = np.meshgrid(f["x_ang_row"], f["y_ang_col"])
    assert ray_trace.ff_fwhm_est(f["beam_fft"], xa, ya) == pytest.approx(float(f["fwhm"]), abs=1e-12)


def test_freq_specs_from_table(feed_table):
    specs = geometry.freq_specs_from_table([90, 106.7, 130], feed_table)
    k, fx, fy = specs[1]
    assert k == 2 * np.pi / geometry.ghz_to_m(106.7)
    row = feed_table[feed_table[:, 0] == 106][0]  # int() truncation, SURVEY appendix B.2
    assert fx == np.deg2rad(row[1]) and fy == np.deg2rad(row[2])
    s2 = geometry.freq_specs_from_table([90], feed_table, fwhm_scale=np.sqrt(2))
    assert s2[0][1] == specs[0][1] * np.sqrt(2)
    with pytest.raises(KeyError):
        geometry.freq_specs_from_table([60], feed_table)
    with pytest.raises(ValueError):
        geometry.make_freqs([])
    with pytest.raises(ValueError):
        geometry.make_freqs([(1, 1, 1)] * 65)


def test_partition_tile_rows_is_a_cyclic_cover():
    for n_scan, world in ((31623, 8), (150, 4), (16, 3), (17, 2), (1, 2)):
        rows = [sdist.partition_tile_rows(n_scan, world, r) for r in range(world)]
        allr = np.sort(np.concatenate(rows))
        assert np.array_equal(allr, np.arange((n_scan + 15) // 16))
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1
        assert all(r.dtype == np.int32 for r in rows)
        for r, rr in enumerate(rows):
            assert np.all(rr % world == r)


def test_weighted_tile_row_deal():
    """Shares of the rays by weighted round-robin: an exact cover, sizes proportional to the
    weights, rows of every rank spread over the whole grid; equal weights = the cyclic deal."""
    n_scan, world = 31623, 8
    w = [0.9] + [1.0] * 7
    rows = [sdist.partition_tile_rows(n_scan, world, r, weights=w) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(1977))
    sizes = np.array([len(r) for r in rows])
    assert abs(sizes[0] / sizes[1:].mean() - 0.9) < 0.01 and sizes[1:].max() - sizes[1:].min() <= 1
    assert all(np.diff(r).max() <= 2 * world for r in rows)          # no rank owns a contiguous block
    for r in range(world):
        assert np.array_equal(sdist.partition_tile_rows(n_scan, world, r, weights=[2.0] * world),
                              sdist.partition_tile_rows(n_scan, world, r))
    with pytest.raises(ValueError):
        sdist.partition_tile_rows(n_scan, 2, 0, weights=[1.0, 0.0])


@pytest.mark.parametrize("n_scan,world", [(60, 1), (60, 2), (150, 8), (31623, 8), (7, 4), (16, 3)])
def test_partition_rows_covers_exactly(n_scan, world):
    blocks = [sdist.partition_rows(n_scan, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n_scan
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 2 * sdist.ROW_ALIGN
    assert all(a % sdist.ROW_ALIGN == 0 for a, _ in blocks if a < n_scan)


def test_partition_items_round_robin():
    parts = [sdist.partition_items(11, 4, r) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(11))
    assert parts[0] == [0, 4, 8] and parts[3] == [3, 7]
    with pytest.raises(ValueError):
        sdist.partition_rows(10, 2, 2)


# ---- world_size-2 gloo run of the ray-sharded near field (tracer replaced by the oracle) ----
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_trace_fn(tele_geo, rx_arr, specs, grid_n, extent_mm, *, opl_ref=0.0, row_begin=0,
                     row_end=None, tile_rows=None, **_):
    """CPU stand-in with the signature of ray_trace.trace_bin_device (test infrastructure)."""
    from oracle import oracle
    n_scan = int(tele_geo.n_scan)
    if tile_rows is not None:   # 16-row tile rows of the launch grid (dist.partition_tile_rows)
        blocks = [(16 * int(t), min(16 * int(t) + 16, n_scan)) for t in tile_rows]
    else:
        blocks = [(row_begin, n_scan if row_end is None else row_end)]
    re = torch.zeros((len(rx_arr), len(specs), grid_n, grid_n), dtype=torch.float32)
    im, cnt = torch.zeros_like(re), torch.zeros((len(rx_arr), grid_n, grid_n), dtype=torch.float32)
    stats = torch.zeros((len(rx_arr), 8), dtype=torch.float64)
    for i, rx in enumerate(rx_arr):
        for b0, b1 in blocks:
            out = oracle.rx_to_lyot(rx, tele_geo, ray_begin=b0 * n_scan, ray_end=b1 * n_scan,
                                    n_threads=1)
            for f, (k, _, _) in enumerate(specs):
                r, m, c, sums = oracle.bin_near_field(out, float(k), grid_n, extent_mm, opl_ref)
                re[i, f] += torch.from_numpy(r).float()
                im[i, f] += torch.from_numpy(m).float()
                if f == 0:
                    cnt[i] += torch.from_numpy(c).float()
            stats[i, 0] += sums["n_valid"]
            stats[i, 1] += sums["sum_opl"]
            stats[i, 7] += sums["n_binned"]
    return re, im, cnt, stats


def _field_fn(re, im, cnt, stats, specs, opl_ref=0.0):
    return torch.complex(re, im), stats.numpy()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        t = oracle.default_tele_geo(n_scan=40, y_source=oracle.Y_LYOT + 300)
        b, cnt, stats = sdist.near_field_binned_distributed(
            t, [[0, 0, 0], [60, 0, -90]], 16, 42.0, trace_fn=_oracle_trace_fn, field_fn=_field_fn)
        sweep = sdist.sweep_distributed(list(range(5)), lambda p: (p, rank))
        # configs[3]: receivers dealt round-robin, per-field scalars merged in receiver order
        rxs = np.array([[10.0 * i, 0.0, -5.0 * i] for i in range(5)])
        calls = []

        def fake_sweep(tele_geo, rx, freqs, grid_n, extent_cm, pad):
            calls.append(rx.copy())
            return {"fwhm_arcmin": rx[:, :1] * np.arange(1, len(freqs) + 1)[None, :],
                    "peak_index": np.repeat(rx[:, None, None, 0], 2, axis=2).repeat(len(freqs), 1).astype(np.int64),
                    "n_valid": rx[:, 2] - 1.0}
        bs = sdist.beam_sweep_distributed(t, rxs, [1, 2, 3], 8, 42.0, 0, sweep_fn=fake_sweep)
        assert len(calls) == 1 and np.array_equal(calls[0], rxs[rank::world])
        assert np.array_equal(bs["fwhm_arcmin"], rxs[:, :1] * np.arange(1, 4)[None, :])
        assert bs["peak_index"].shape == (5, 3, 2) and bs["peak_index"].dtype == np.int64
        assert np.array_equal(bs["peak_index"][:, 0, 0], rxs[:, 0].astype(np.int64))
        assert np.array_equal(bs["n_valid"], rxs[:, 2] - 1.0)
        # one host array shared by the ranks (POSIX shared memory): every rank fills its slab
        sh = sdist.SharedHostArray((6, 5), np.complex64)
        c0, c1 = sdist.slab_cells(30, world, rank)
        sh.array[c0:c1] = np.arange(c0, c1) * (1 + 1j)
        dist.barrier()
        assert np.array_equal(sh.view(), (np.arange(30) * (1 + 1j)).astype(np.complex64).reshape(6, 5))
        sh.close()
        # more ranks than receivers: the idle rank still takes part in the gather
        one = sdist.beam_sweep_distributed(t, rxs[:1], [1], 8, 42.0, 0, sweep_fn=fake_sweep)
        assert one["fwhm_arcmin"].shape == (1, 1) and one["n_valid"][0] == -1.0
        # the single-buffer form bench.py uses: planes are views of one flat tensor, one collective
        re, im, c2, flat = sdist.alloc_planes(2, 1, 16, device="cpu")
        st = torch.zeros((2, 8), dtype=torch.float64)
        r0, r1 = sdist.partition_rows(40, world, rank)
        a, bb, cc, ss = _oracle_trace_fn(t, np.array([[0., 0, 0], [60, 0, -90]]),
                                         [(t.k, t.th_fwhp_x, t.th_fwhp_y)], 16, 420.0,
                                         row_begin=r0, row_end=r1)
        re.copy_(a); im.copy_(bb); c2.copy_(cc); st.copy_(ss)
        sdist.reduce_bins(re, im, c2, st, flat=flat)
        assert flat.data_ptr() == re.data_ptr() and torch.equal(c2, cnt)
        assert (torch.complex(re, im) - b).abs().max().item() < 1e-6
        q.put((rank, b.numpy(), cnt.numpy(), stats, sweep))
    finally:
        dist.destroy_process_group()


def test_ray_sharded_near_field_world2_gloo():
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # single-process answer
    t = oracle.default_tele_geo(n_scan=40, y_source=oracle.Y_LYOT + 300)
    re, im, cnt, stats = _oracle_trace_fn(t, np.array([[0., 0, 0], [60, 0, -90]]),
                                          [(t.k, t.th_fwhp_x, t.th_fwhp_y)], 16, 420.0,
                                          row_begin=0, row_end=40)
    for rank, b, c, s, sweep in res:
        assert np.array_equal(c, cnt.numpy())
        assert np.abs(b - torch.complex(re, im).numpy()).max() < 1e-4
        assert np.allclose(s[:, [0, 7]], stats.numpy()[:, [0, 7]])
        assert np.allclose(s[:, 1], stats.numpy()[:, 1], rtol=1e-12)
        assert sweep == [(0, 0), (1, 1), (2, 0), (3, 1), (4, 0)]


def test_spline_operator_reproduces_the_reference_sim_data_golden():
    """Row f1: ``sim_data``'s bicubic spline resampling is the separable linear map My z Mx^T of
    the grid values; the operators ``ray_trace._cubic_spline_operator`` builds (FITPACK's
    not-a-knot knots, clamped arguments) applied in numpy to the 100 x 100 grid the reference's
    sim2d produced give the reference's stage-grid arrays -- the arithmetic the device runs."""
    g = golden("farfield_cfg2_240.npz")
    a, p = g["sb2_a"], g["sb2_p"]
    x_data = np.arange(-40.0, 40.0, 2)
    Mx = ray_trace._cubic_spline_operator(g["sb2_x"][:, 0], x_data)
    My = ray_trace._cubic_spline_operator(g["sb2_y"][0, :], x_data)
    beam = a.T * np.exp(1j * p.T)
    amp = My @ (np.abs(beam) / np.abs(beam).max()) @ Mx.T
    ph = My @ np.arctan2(beam.imag, beam.real) @ Mx.T
    assert np.abs(amp / amp.max() - g["a_data"]).max() < 1e-12
    assert np.abs(ph - g["p_data"]).max() < 1e-12


def test_spline_operator_matches_fitpack_on_an_irregular_grid():
    from scipy.interpolate import RectBivariateSpline
    rng = np.random.default_rng(4)
    x = np.sort(rng.uniform(-3, 5, 37))
    y = np.sort(rng.uniform(0, 2, 11))
    z = rng.standard_normal((len(y), len(x)))
    xn = np.linspace(-4, 6, 53)     # beyond both ends: clamped like fpbisp
    yn = np.linspace(-0.5, 2.5, 29)
    want = RectBivariateSpline(x, y, z.T, kx=3, ky=3, s=0)(xn, yn).T
    got = ray_trace._cubic_spline_operator(y, yn) @ z @ ray_trace._cubic_spline_operator(x, xn).T
    assert np.abs(got - want).max() < 1e-11 * np.abs(want).max()
    with pytest.raises(ValueError):
        ray_trace._cubic_spline_operator([0.0, 1.0, 1.0, 2.0], [0.5])



def test_peer_plane_layout_and_cell_slabs():
    """Host logic of the fused peer-memory reduce (dist.PeerPlanes / csrc/peer.cu): the buffer
    layout is aligned and disjoint, and the cell slabs tile a field exactly with every slab
    starting on a multiple of 4 cells (the kernel's 16-byte loads)."""
    from sosat_optics_b200 import dist as sdist
    for n_rx, n_freq, n in ((1, 1, 2048), (2, 3, 66), (3, 1, 7)):
        L = sdist.peer_layout(n_rx, n_freq, n)
        n_ri, n_c = n_rx * n_freq * n * n, n_rx * n * n
        regions = [("re", 4 * n_ri), ("im", 4 * n_ri), ("cnt", 4 * n_c), ("stats", 64 * n_rx),
                   ("stats_sum", 64 * n_rx), ("field", 8 * n_ri)]
        end = 0
        for name, nbytes in regions:
            assert L[name] % 256 == 0 and L[name] >= end, name
            end = L[name] + nbytes
        assert L["bytes"] >= end and L["accum_bytes"] == L["stats_sum"]
        for world in (1, 2, 3, 8, 16):
            slabs = [sdist.slab_cells(n * n, world, r) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == n * n
            for (a0, a1), (b0, b1) in zip(slabs, slabs[1:]):
                assert a1 == b0
            assert all((b % 4 == 0 and e > b) or e == b for b, e in slabs)  # empty slabs allowed
            sizes = [e - b for b, e in slabs if e > b]
            assert max(sizes) - min(sizes) < 8 or len(sizes) < world  # one 4-cell unit + ragged tail


def test_ot_geo_surface_functions_match_the_reference():
    """``from ot_geo import *`` still provides the reference's Python-level sag / slope / frame
    functions (ot_geo.py:49-458); values of the reference itself at seeded points, incl. NaN
    beyond the conic domain of lens 1b (golden: oracle/make_golden.py ot_geo)."""
    from sosat_optics_b200 import ot_geo
    g = golden("ot_geo_surfaces.npz")
    x, y, z = g["x"], g["y"], g["z"]
    with np.errstate(invalid="ignore"):
        for name in ("1a", "1b", "2a", "2b", "3a", "3b"):
            sag = getattr(ot_geo, f"z{name}")(x, y)
            assert np.array_equal(np.isnan(sag), np.isnan(g[f"z{name}"]))
            assert np.allclose(sag, g[f"z{name}"], rtol=1e-14, atol=0, equal_nan=True)
            gx, gy = getattr(ot_geo, f"d_z{name}")(x, y)
            assert np.allclose(gx, g[f"d_z{name}_x"], rtol=1e-13, atol=0, equal_nan=True)
            assert np.allclose(gy, g[f"d_z{name}_y"], rtol=1e-13, atol=0, equal_nan=True)
            y_keep = y.copy()
            fwd = np.array(getattr(ot_geo, f"tele_into_m{name}")(x, y, z))
            assert np.array_equal(y, y_keep)                     # arguments are left alone
            assert np.allclose(fwd, g[f"tele_into_m{name}"], rtol=1e-14, atol=1e-12)
            back = np.array(getattr(ot_geo, f"m{name}_into_tele")(x, y, z))
            assert np.allclose(back, g[f"m{name}_into_tele"], rtol=1e-14, atol=1e-12)
            rt = getattr(ot_geo, f"m{name}_into_tele")(*getattr(ot_geo, f"tele_into_m{name}")(x, y, z))
            assert np.allclose(np.array(rt), np.array([x, y, z]), atol=1e-10)
    assert "1.5 * 2" not in ot_geo.z1b.__doc__ and ot_geo.z1b.__name__ == "z1b"
