"""The compiled CPU port (oracle/cpu_port) against the numpy oracle: same operators, same RTR
trajectory.  It is the 'reference CPU' column of bench.py, so it has to be right too."""
import numpy as np
import pytest

from oracle import pgo
from oracle.cpu_port import CpuProblem


@pytest.mark.parametrize("name,r", [("tinyGrid3D", 5), ("smallGrid3D", 5), ("smallGrid3D", 3), ("sphere2500", 5)])
def test_cpu_port_matches_numpy_oracle(datasets, name, r):
    meas, n, z = datasets(name)
    d = meas.d
    Q = pgo.connection_laplacian(meas, n)
    rng = np.random.default_rng(0)
    G = rng.standard_normal((r, (d + 1) * n)) if name != "sphere2500" else np.zeros((r, (d + 1) * n))
    op = pgo.QuadraticProblem(Q, G, d)
    cp = CpuProblem(Q, G, d)
    X = pgo.manifold_project(rng.standard_normal((r, (d + 1) * n)), d)
    V = rng.standard_normal(X.shape)
    assert np.linalg.norm(cp.qx(X) - op.XQ(X)) <= 1e-12 * np.linalg.norm(op.XQ(X))
    cp.factorize()
    Z = op.precon_solve(V)
    assert np.linalg.norm(cp.solve(V) - Z) <= 1e-8 * np.linalg.norm(Z)
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    Xc, rc = cp.optimize(X0)
    Xo, ro = pgo.optimize(op, X0)
    assert (rc["outer"], rc["inner"], rc["accepted"]) == (ro.outer, ro.inner_total, ro.accepted)
    assert abs(rc["f_init"] - ro.fInit) <= 1e-12 * abs(ro.fInit)
    assert abs(rc["f_opt"] - ro.fOpt) <= 1e-9 * abs(ro.fOpt)
    assert np.linalg.norm(Xc - Xo) <= 1e-6 * np.linalg.norm(Xo)


def test_cpu_port_2d(datasets):
    meas, n, z = datasets("city10000")
    keep = np.where((meas.p1 < 400) & (meas.p2 < 400))[0]
    sub = meas.subset(keep)
    n, d, r = 400, 2, 3
    Q = pgo.connection_laplacian(sub, n)
    G = np.zeros((r, 3 * n))
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"][:, :3 * n]
    Xc, rc = CpuProblem(Q, G, d).optimize(X0)
    Xo, ro = pgo.optimize(pgo.QuadraticProblem(Q, G, d), X0)
    assert (rc["outer"], rc["inner"]) == (ro.outer, ro.inner_total)
    # the third outer iteration runs the full 50 tCG steps (status MAXITER) on a problem with
    # cond ~ 1e4: 50 Krylov steps amplify the 1e-14 operator differences to ~1e-5 on the iterate
    # (operators themselves are compared to 1e-10 below)
    assert abs(rc["f_opt"] - ro.fOpt) <= 1e-4 * abs(ro.fOpt)
    assert np.linalg.norm(Xc - Xo) <= 1e-4 * np.linalg.norm(Xo)
    cp = CpuProblem(Q, G, d)
    cp.factorize()
    V = np.random.default_rng(1).standard_normal((r, 3 * n))
    Z = cp.solve(V)
    assert np.linalg.norm((Q + 0.1 * np.eye(3 * n)) @ Z.T - V.T) <= 1e-10 * np.linalg.norm(V)
