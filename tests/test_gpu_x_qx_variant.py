"""GPU test of the prefetching measurement variant of the stand-alone Q*X (dpgo_set_qx_variant(h, 1, .)):
same arithmetic as the default kernel plus cp.async.bulk.prefetch.L2 hints, so the product must be
identical.  Written after this round's GPU budget was spent (no default path uses the variant): outside
DPGO_B200_EXPERIMENTAL=1 a failure is an expected failure.  Last in collection order."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
STRICT = os.environ.get("DPGO_B200_EXPERIMENTAL") == "1"


@pytest.mark.xfail(condition=not STRICT, reason="Q*X prefetch variant: first device run pending", strict=False)
@pytest.mark.parametrize("name,r", [("smallGrid3D", 5), ("sphere2500", 5), ("city10000", 3)])
def test_qx_prefetch_variant_is_identical(datasets, name, r):
    import dpgo_b200
    meas, n, _ = datasets(name)
    d = meas.d
    gp = dpgo_b200.problem_from_measurements(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r,
                                             build_precon=False)
    X = np.random.default_rng(5).standard_normal((r, (d + 1) * n))
    ref = gp.qx(X)
    for dist in (0, 8, 1000, 10 * n):
        gp.set_qx_variant(1, dist)
        assert np.array_equal(gp.qx(X), ref), dist
    gp.set_qx_variant(0, 0)
    assert np.array_equal(gp.qx(X), ref)
    gp.close()
