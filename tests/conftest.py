import os
import sys

# Single-threaded BLAS for the whole suite, set before numpy loads: a multi-threaded OpenBLAS call made after
# tests/test_exchange_gloo.py has forked its multiprocessing manager can deadlock in this process (seen as
# np.linalg.inv never returning); nothing here is large enough to need the threads.
for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402
import pytest

try:    # numpy may have been imported by a pytest plugin before this file: cap the live pools as well
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1)
except Exception:   # pragma: no cover
    pass

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_dataset(name):
    """Fixture written by tools/make_fixtures.py from the reference's data/<name>.g2o."""
    from oracle import pgo
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d, n = int(z["d"]), int(z["n"])
    meas = pgo.make_measurements(d, z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"])
    return meas, n, z


@pytest.fixture(scope="session")
def datasets():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_dataset(name)
        return cache[name]
    return get
