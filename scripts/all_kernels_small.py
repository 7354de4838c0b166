"""Small end-to-end run that launches every kernel of the library once (bring-up / smoke script;
compute-sanitizer is closed on the GPU pool, so bounds are covered by the parity tests instead)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from sosat_optics_b200 import geometry, opt_analyze as oa, ot_geo, ray_trace as rt  # noqa: E402

t = ot_geo.SatGeo()
t.n_scan = 70
t.y_source = ot_geo.y_lyot + 300
t.lambda_ = oa.ghz_to_m(90)
t.k = 2 * np.pi / t.lambda_
out = rt.rx_to_lyot([10, 0, -20], t)
sb = rt.getNearField(t, [0, 0, 0])
m = rt.near_field_metrics(t, [0, 0, 0])
table = np.load(os.path.join(ROOT, "feedhorn", "fwhm_feedhorns.npy"))
specs = geometry.freq_specs_from_table([80.0, 90.0, 100.0], table)
bf = rt.near_field_binned(t, [[0, 0, 0], [100, 0, 60]], grid_n=40, extent_cm=42.0, freqs=specs)
for n in (100, 256, 384):
    rng = np.random.default_rng(n)
    b = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64)
    B = oa.far_field(b)
    ref = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(b.astype(np.complex128))))
    assert np.abs(B - ref).max() < 1e-4 * np.abs(ref).max()
    bc, ph = oa.a2b(np.abs(b).astype(np.float64), np.angle(b, deg=True).astype(np.float64))
    af, ap = oa.b2a(np.abs(b).astype(np.float64), np.angle(b, deg=True).astype(np.float64))
x = (np.arange(64) - 32) * 0.5
X, Y = np.meshgrid(x, x)
_, _, conv = oa.beam_convolve(X, Y, np.exp(-(X ** 2 + Y ** 2) / 50).astype(np.complex128), 2.0, 3.0, 0)
th = np.linspace(-0.02, 0.02, 33)
Bz = oa.far_field(np.exp(-(X ** 2 + Y ** 2) / 50).astype(np.complex64), 0.5, 3.33e-3, theta_x=th, theta_y=th)
xa, ya = oa.coords_spat_to_ang(X / 1e2, Y / 1e2, 90.0)
print("fwhm", rt.ff_fwhm_device(oa.far_field(np.exp(-(X ** 2 + Y ** 2) / 50).astype(np.complex64)), xa, ya))
pts = np.random.default_rng(1).uniform(-0.2, 0.2, (2, 500))
Bi = oa.far_field_irregular(pts[0], pts[1], np.ones(500, complex), th, th, 3.33e-3)
print("all_kernels_small ok", int((out[4] != 0).sum()), len(sb.a_sim), m["n_valid"], bf.b.shape)
