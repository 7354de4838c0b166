import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def feed_table():
    return golden("feedhorn_fwhm.npz")["table"]


def tele_geo_from_golden(g, n_scan=None):
    """A SatGeo configured like the reference run that produced golden file ``g``."""
    from sosat_optics_b200 import ot_geo
    t = ot_geo.SatGeo()
    t.n_scan = int(g["n_scan"]) if n_scan is None else n_scan
    t.y_source = float(g["y_source"])
    t.lambda_ = float(g["lambda_"])
    t.k = float(g["k"])
    t.th_fwhp_x = g["th_fwhp_x"]
    t.th_fwhp_y = g["th_fwhp_y"]
    return t


def phase_diff(p, q):
    """|p - q| modulo 2 pi."""
    d = np.mod(np.asarray(p) - np.asarray(q) + np.pi, 2 * np.pi) - np.pi
    return np.abs(d)
