"""CPU tests of the robust re-weighting path: the oracle's restatement of
computeMeasurementError / RobustCost / updateMeasurementWeights and the product-side scalar
mirror dpgo_b200/robust.py (no GPU needed)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import pgo, rbcd as orbcd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _robust_module():
    # by path: importing the dpgo_b200 package loads the CUDA library
    spec = importlib.util.spec_from_file_location("_robust", os.path.join(ROOT, "dpgo_b200", "robust.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_measurement_errors_sum_to_the_cost(datasets):
    """sum_e w_e * computeMeasurementError_e == 2 f(X) = <X Q, X>: pins the per-edge error
    (src/DPGO_utils.cpp:501-507) against the connection Laplacian (:272-344)."""
    meas, n, z = datasets("smallGrid3D")
    d, r = meas.d, 5
    rng = np.random.default_rng(3)
    X = rng.standard_normal((r, (d + 1) * n))
    w = rng.uniform(0.1, 1.0, len(meas))
    meas_w = pgo.make_measurements(d, meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, weight=w)
    Q = pgo.connection_laplacian(meas_w, n)
    two_f = float(np.sum((X @ Q) * X))
    total = 0.0
    for k in range(len(meas)):
        T1 = X[:, meas.p1[k] * (d + 1):(meas.p1[k] + 1) * (d + 1)]
        T2 = X[:, meas.p2[k] * (d + 1):(meas.p2[k] + 1) * (d + 1)]
        total += w[k] * pgo.measurement_error(meas.R[k], meas.t[k], meas.kappa[k], meas.tau[k],
                                              T1[:, :d], T1[:, d], T2[:, :d], T2[:, d])
    assert abs(total - two_f) <= 1e-10 * abs(two_f)


@pytest.mark.parametrize("kind", ["L2", "L1", "TLS", "Huber", "GM", "GNC_TLS"])
def test_robust_cost_mirror_matches_oracle(kind):
    rb = _robust_module()
    a, b = rb.RobustCost(kind), pgo.RobustCost(kind)
    r = np.concatenate([np.geomspace(1e-3, 1e4, 200), [3.0, 10.0, 5.0]])
    for step in range(25):          # beyond GNCMaxNumIters = 20: mu stops growing
        wa = a.weights(r)
        wb = np.array([b.weight(x) for x in r])
        assert np.allclose(wa, wb, rtol=1e-14, atol=0)
        assert abs(a.mu - b.mu) <= 1e-18
        a.update(); b.update()
    if kind == "GNC_TLS":
        assert abs(b.mu - 1e-4 * 1.4 ** 20) < 1e-12
        # eq. (14) of the GNC paper: 1 below mu/(mu+1) c^2, 0 above (mu+1)/mu c^2, monotone between
        c = pgo.RobustCost("GNC_TLS"); c.mu = 0.5
        assert c.weight(np.sqrt(0.5 / 1.5 * 25) * 0.999) == 1.0
        assert c.weight(np.sqrt(1.5 / 0.5 * 25) * 1.001) == 0.0
        mids = [c.weight(x) for x in np.linspace(3.0, 8.5, 20)]
        assert all(0 <= y <= 1 for y in mids) and all(x >= y for x, y in zip(mids, mids[1:]))


def test_chi2_threshold():
    from scipy.stats import chi2
    rb = _robust_module()
    for q in (0.5, 0.9, 0.99):
        assert abs(rb.chi2_threshold_3d(q) - np.sqrt(chi2.ppf(q, 6))) < 1e-9
    assert rb.chi2_threshold_3d(1.0) == 1e5


def test_partition_fixed_flags_and_outlier_rejection(datasets):
    """updateMeasurementWeights (src/PGOAgent.cpp:1104-1142): odometry edges (consecutive global
    ids, also across the partition boundary) keep weight 1; a corrupted loop closure is driven
    to weight 0 by the GNC schedule while the inliers return to 1."""
    meas, n, z = datasets("smallGrid3D")
    d, r, A = meas.d, 5, 5
    t_bad = meas.t.copy()
    lc = np.where(meas.p2 != meas.p1 + 1)[0]
    bad = lc[7]
    t_bad[bad] += 25.0                          # gross translation outlier on one loop closure
    mb = pgo.make_measurements(d, meas.p1, meas.p2, meas.R, t_bad, meas.kappa, meas.tau)
    team = orbcd.Team(mb, n, A, r, acceleration=False)
    ranges, parts = orbcd.partition(mb, n, A)
    nfixed = sum(int(P.fixed.sum()) + int(S.fixed.sum()) // 1 for P, S in parts)
    # every odometry edge is private to one robot or shared by exactly two
    n_odo = int(np.sum(meas.p2 == meas.p1 + 1))
    shared_fixed = sum(int(S.fixed.sum()) for _, S in parts)
    assert nfixed - shared_fixed // 2 == n_odo and shared_fixed == 2 * (A - 1)
    team.set_X(pgo.lifting_matrix(d, r) @ z["T_chordal"])
    colors = orbcd.robot_graph_coloring(team.agents)
    team.robust = [pgo.RobustCost("GNC_TLS", gnc_barc=5.0, gnc_init_mu=1e-2, gnc_mu_step=2.0) for _ in team.agents]
    k = 0
    for outer in range(12):
        for _ in range(4):
            team.step_colored(colors, k); k += 1
        team.update_weights()
    w_bad, w_in = [], []
    for a, (P, S) in zip(team.agents, parts):
        for M, src in ((a.graph.private, P), (a.graph.shared, S)):
            for j in range(len(M)):
                if M.fixed[j]:
                    assert M.weight[j] == 1.0
                elif abs(M.t[j] - src.t[j]).max() == 0 and np.abs(M.t[j]).max() > 20:
                    w_bad.append(M.weight[j])
                else:
                    w_in.append(M.weight[j])
    assert len(w_bad) >= 1 and max(w_bad) == 0.0
    # residuals of inliers are chi-square distributed with 6 dof: P(r > 5) = 3e-4
    assert np.mean(np.array(w_in) == 1.0) > 0.99
