import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product libraries and the oracle are built (no-op when up to date)."""
    from hitl_slam_b200 import build
    build.build_all()
    from oracle import pyoracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        pyoracle.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def host():
    from hitl_slam_b200 import HostLib
    return HostLib()


@pytest.fixture(scope="session")
def gpu():
    from hitl_slam_b200 import HitlGpu
    return HitlGpu(0)


_MAPS = {}


def synth_map(name, **kw):
    from hitl_slam_b200 import synth
    key = (name, tuple(sorted(kw.items())))
    if key not in _MAPS:
        _MAPS[key] = synth.generate(name, **kw)
    return _MAPS[key]


@pytest.fixture(scope="session")
def maps():
    return synth_map


def random_scans(rng, n_scans, lo, hi, ties=True, empty=()):
    """Ragged random scans with duplicated coordinates (tie cases of the tree builder)."""
    sizes = [0 if i in empty else int(rng.integers(lo, hi + 1)) for i in range(n_scans)]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    m = int(off[-1])
    pts = (rng.normal(size=(m, 2)) * 2.0).astype(np.float32)
    if ties and m:
        pts[::7, 0] = np.round(pts[::7, 0], 1)
        pts[::5, 1] = np.round(pts[::5, 1], 1)
    ang = rng.uniform(0, 2 * np.pi, m)
    nrm = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)
    return off, pts, nrm


def assert_same_stf(a, b):
    for key in ("pair_i", "pair_j", "pair_off", "k", "idx"):
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
