import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture
def portable_trig():
    """Switch the oracle to the portable cos / sin (the routine the kernels use) for one test: GPU, host
    emulation and oracle are then bit-identical over whole episodes."""
    from oracle import cpo
    cpo.set_portable_trig(True)
    try:
        yield
    finally:
        cpo.set_portable_trig(False)


FIXTURES = ["c10_t0", "c10_t7", "c30_t7", "c50_t0"]


def fixture_conc(name):
    return int(name[1:3]) / 100.0


@pytest.fixture(scope="session")
def golden_layouts():
    from benchpush_b200.layouts import load_packed
    return {n: load_packed(os.path.join(GOLD, "ice_%s.npz" % n)) for n in FIXTURES}


@pytest.fixture(scope="session")
def pins():
    return np.load(os.path.join(GOLD, "shipice_oracle_pins.npz"))


@pytest.fixture(scope="session")
def spfa_golden():
    z = np.load(os.path.join(GOLD, "spfa_golden.npz"))
    cases, off = [], 0
    for (H, W), src in zip(z["shapes"], z["sources"]):
        n = int(H) * int(W)
        cases.append(dict(map=z["maps"][off:off + n].reshape(H, W), source=(int(src[0]), int(src[1])),
                          dists=z["dists"][off:off + n].reshape(H, W), parents=z["parents"][off:off + n].reshape(H, W)))
        off += n
    return cases


SHIP_VERTS = [[1., -0.], [0.9, 0.10], [0.5, 0.25], [0.25, 0.25], [0, 0.25], [-0.25, 0.25], [-0.5, 0.25], [-0.75, 0.25],
              [-1., 0.25], [-1., -0.25], [-0.75, -0.25], [-0.5, -0.25], [-0.25, -0.25], [0, -0.25], [0.25, -0.25],
              [0.5, -0.25], [0.9, -0.10]]
