"""CPU tests: the oracle against the reference's own known-answer tests and against the
committed golden fixtures (no GPU)."""
import numpy as np
import pytest

from oracle import pgo, rbcd


def _triangle():
    # tests/testTriangleGraph.cpp:15-47
    Tw0 = np.eye(4)
    Tw1 = np.array([[0.1436, 0.7406, 0.6564, 1], [-0.8179, -0.2845, 0.5000, 1],
                    [0.5571, -0.6087, 0.5649, 1], [0, 0, 0, 1.0]])
    Tw2 = np.array([[-0.4069, -0.4150, -0.8138, 2], [0.4049, 0.7166, -0.5679, 2],
                    [0.8188, -0.5606, -0.1236, 2], [0, 0, 0, 1.0]])
    Ttrue = np.hstack([Tw0[:3], Tw1[:3], Tw2[:3]])

    def rel(A, B):
        dT = np.linalg.inv(A) @ B
        return dT[:3, :3], dT[:3, 3]
    es = [(0, 1, *rel(Tw0, Tw1)), (1, 2, *rel(Tw1, Tw2)), (0, 2, *rel(Tw0, Tw2))]
    meas = pgo.make_measurements(3, [e[0] for e in es], [e[1] for e in es], [e[2] for e in es],
                                 [e[3] for e in es], [1.0] * 3, [1.0] * 3)
    return meas, Ttrue


def test_triangle_known_answer():
    """tests/testTriangleGraph.cpp:56-61: solvePGO (chordal init + RTR, r = d) equals Ttrue to 1e-4."""
    meas, Ttrue = _triangle()
    T0 = pgo.chordal_initialization(meas, 3)
    prob = pgo.QuadraticProblem(pgo.connection_laplacian(meas, 3), np.zeros((3, 12)), 3)
    Y, res = pgo.optimize(prob, T0)
    assert np.linalg.norm(pgo.round_solution(Y, 3) - Ttrue) <= 1e-4
    # and from a perturbed start, so that RTR/tCG actually run
    rng = np.random.default_rng(0)
    Y0 = pgo.manifold_project(Ttrue + 0.3 * rng.standard_normal(Ttrue.shape), 3)
    Y, res = pgo.optimize(prob, Y0, pgo.ROptParameters(RTR_iterations=50, RTR_tCG_iterations=500,
                                                       gradnorm_tol=1e-8))
    assert res.inner_total > 0 and res.gradNormOpt < 1e-8
    assert np.linalg.norm(pgo.round_solution(Y, 3) - Ttrue) <= 1e-4


def test_prior_known_answer():
    """tests/testPGO.cpp:131-190: two poses + prior on pose 1 -> both poses on the prior (1e-6)."""
    m = pgo.make_measurements(3, [0], [1], [np.eye(3)], [np.zeros(3)], [10000.0], [100.0])
    T = pgo.odometry_initialization(m, 2)
    pr = pgo.project_rotation(np.array([[0.7236, 0.1817, 0.6658], [-0.6100, 0.6198, 0.4938],
                                        [-0.3230, -0.7634, 0.5594]]))
    P = np.hstack([pr, np.zeros((3, 1))])
    assert np.linalg.norm(T[:, :4] - P) > 1e-6
    g = pgo.LocalGraph(0, 3, 3, 2, m, None, {1: P})
    prob = pgo.QuadraticProblem(pgo.construct_Q(g), pgo.construct_G(g, {}), 3)
    Y, res = pgo.optimize(prob, T, pgo.ROptParameters(RTR_iterations=50, RTR_tCG_iterations=500,
                                                      gradnorm_tol=1e-5))
    assert np.linalg.norm(Y[:, :4] - P) < 1e-6
    assert np.linalg.norm(Y[:, 4:] - P) < 1e-6


def test_projection_properties():
    """tests/testUtils.cpp:13-54: projections give Y^T Y = I to 1e-5 (r=5, d=3, n=100)."""
    rng = np.random.default_rng(1)
    M = rng.standard_normal((5, 4 * 100))
    X = pgo.manifold_project(M, 3)
    for i in range(100):
        Y = X[:, 4 * i:4 * i + 3]
        assert np.linalg.norm(Y.T @ Y - np.eye(3)) < 1e-5
        assert np.array_equal(X[:, 4 * i + 3], M[:, 4 * i + 3])
    L = pgo.lifting_matrix(3, 5)
    assert np.linalg.norm(L.T @ L - np.eye(3)) < 1e-12
    assert np.array_equal(L, pgo.lifting_matrix(3, 5))
    # QF retraction stays on the manifold and is a retraction (R(0) = id)
    V = pgo.tangent_project(X, rng.standard_normal(X.shape), 3)
    X2 = pgo.retract_qf(X, 0.3 * V, 3)
    for i in range(100):
        Y = X2[:, 4 * i:4 * i + 3]
        assert np.linalg.norm(Y.T @ Y - np.eye(3)) < 1e-12
    assert np.linalg.norm(pgo.retract_qf(X, 0 * V, 3) - X) < 1e-12


def test_derivatives_consistent():
    """Finite differences pin egrad / rhess against f (guards the restated ROPTLIB formulas)."""
    rng = np.random.default_rng(2)
    n, d, r = 12, 3, 5
    p1 = list(range(n - 1)) + [0, 3, 5]
    p2 = list(range(1, n)) + [7, 9, 11]
    m = len(p1)
    Rs = [pgo.project_rotation(rng.standard_normal((3, 3))) for _ in range(m)]
    meas = pgo.make_measurements(3, p1, p2, Rs, rng.standard_normal((m, 3)),
                                 rng.uniform(1, 5, m), rng.uniform(1, 5, m))
    Q = pgo.connection_laplacian(meas, n)
    assert abs(Q - Q.T).max() < 1e-12
    G = rng.standard_normal((r, 4 * n))
    prob = pgo.QuadraticProblem(Q, G, d)
    X = pgo.manifold_project(rng.standard_normal((r, 4 * n)), d)
    V = pgo.tangent_project(X, rng.standard_normal(X.shape), d)
    EG = prob.egrad(X)
    g = pgo.tangent_project(X, EG, d)
    h = 1e-5
    fd = (prob.f(pgo.retract_qf(X, h * V, d)) - prob.f(pgo.retract_qf(X, -h * V, d))) / (2 * h)
    assert abs(fd - np.sum(g * V)) < 1e-6 * max(1, abs(fd))
    # second-order: f(R(hV)) = f + h<g,V> + h^2/2 <V,HV> + O(h^3) along a SECOND-order
    # retraction (the polar one; QF is only first order so it cannot be used here)
    HV = prob.rhess(X, EG, V)
    f0 = prob.f(X)
    h = 1e-3
    polar = lambda s: pgo.manifold_project(X + s * V, d)
    f2 = (prob.f(polar(h)) + prob.f(polar(-h)) - 2 * f0) / h ** 2
    assert abs(f2 - np.sum(V * HV)) < 1e-4 * max(1, abs(f2))
    # preconditioner is the exact inverse of Q + 0.1 I followed by the projection
    Z = prob.precon_solve(V)
    assert np.linalg.norm((Q + 0.1 * np.eye(Q.shape[0])) @ Z.T - V.T) < 1e-9 * np.linalg.norm(V)


@pytest.mark.parametrize("name,cost_tol", [("tinyGrid3D", 1e-9), ("smallGrid3D", 1e-9)])
def test_fixture_golden_costs(datasets, name, cost_tol):
    """The fixture's golden scalars reproduce (guards oracle edits)."""
    meas, n, z = datasets(name)
    d = meas.d
    prob = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((d, (d + 1) * n)), d)
    assert prob.Q.nnz == int(z["q_scalar_nnz"])
    assert abs(2 * prob.f(z["T_chordal"]) - float(z["cost2_chordal"])) <= cost_tol * float(z["cost2_chordal"])
    od = np.where(meas.p2 == meas.p1 + 1)[0]
    Tod = pgo.odometry_initialization(meas.subset(od), n)
    assert abs(2 * prob.f(Tod) - float(z["cost2_odometry"])) <= cost_tol * float(z["cost2_odometry"])
    Tch = pgo.chordal_initialization(meas, n)
    assert np.linalg.norm(Tch - z["T_chordal"]) < 1e-8 * np.linalg.norm(Tch)


def test_sphere2500_converges_to_known_optimum(datasets):
    """sphere2500: RTR from the lifted chordal init reaches the SE-Sync optimum 2f* = 1687.0058."""
    meas, n, z = datasets("sphere2500")
    d, r = 3, 5
    prob = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((r, 4 * n)), d)
    X = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    for _ in range(2):
        X, res = pgo.optimize(prob, X)
    assert abs(2 * res.fOpt - 1687.0058205) < 1e-5
    assert res.gradNormOpt < 1e-2


def test_multi_agent_rbcd_decreases_cost(datasets):
    """Config 1 (smallGrid3D, 5 agents, r = 5, accelerated synchronous RBCD as
    examples/MultiRobotExample.cpp runs it): the cost sequence decreases to the optimum."""
    meas, n, z = datasets("smallGrid3D")
    team = rbcd.Team(meas, n, 5, 5, acceleration=True)
    team.set_X(pgo.lifting_matrix(3, 5) @ z["T_chordal"])
    costs = [team.step_greedy()["cost"] for _ in range(60)]
    assert costs[-1] < costs[0]
    # centralized optimum for comparison
    prob = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((5, 4 * n)), 3)
    X = pgo.lifting_matrix(3, 5) @ z["T_chordal"]
    for _ in range(5):
        X, res = pgo.optimize(prob, X, pgo.ROptParameters(RTR_iterations=10, gradnorm_tol=1e-6))
    assert costs[-1] >= 2 * res.fOpt - 1e-9
    assert (costs[-1] - 2 * res.fOpt) / (2 * res.fOpt) < 1e-4
    # coloured parallel schedule reaches the same optimum
    team2 = rbcd.Team(meas, n, 5, 5, acceleration=True)
    team2.set_X(pgo.lifting_matrix(3, 5) @ z["T_chordal"])
    colors = rbcd.robot_graph_coloring(team2.agents)
    assert len(colors) == 2
    c2 = [team2.step_colored(colors, k)["cost"] for k in range(40)]
    assert 0 <= (c2[-1] - 2 * res.fOpt) / (2 * res.fOpt) < 1e-5


def test_synthetic_grid_generator_matches_the_measurement_spec():
    """dpgo_b200/synthetic.py builds the roofline-scale input of SURVEY 8(d): an L^3 lattice visited
    along a snake path (odometry chain), every other lattice edge a loop closure, kappa = 200,
    tau = 100.  Without noise the ground truth has zero cost under the oracle's Q."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_syn", os.path.join(root, "dpgo_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    L = 6
    g = syn.grid3d(L, seed=3, noise_rot=0.0, noise_t=0.0)
    n = L ** 3
    assert g["n"] == n and g["d"] == 3
    m = 3 * L * L * (L - 1)                                   # lattice edges
    assert len(g["p1"]) == m and np.all(g["p1"] < g["p2"])
    odo = g["p2"] == g["p1"] + 1
    assert odo.sum() == n - 1 and np.all(odo[:n - 1]) and not odo[n - 1:].any()   # chain first, as in the datasets
    assert np.all(g["kappa"] == 200.0) and np.all(g["tau"] == 100.0)
    assert np.allclose(np.einsum("mab,mcb->mac", g["R"], g["R"]), np.eye(3), atol=1e-12)
    meas = pgo.make_measurements(3, g["p1"], g["p2"], g["R"], g["t"], g["kappa"], g["tau"])
    Q = pgo.connection_laplacian(meas, n)
    T = g["T_true"]
    assert abs(np.sum((T @ Q) * T)) < 1e-9                    # consistent measurements: zero cost at the truth
    deg = np.bincount(np.concatenate([g["p1"], g["p2"]]), minlength=n)
    assert deg.max() == 6 and deg.min() == 3                  # 6-neighbour lattice
    gn = syn.grid3d(L, seed=3)                                # with the default noise the cost is small but not zero
    measn = pgo.make_measurements(3, gn["p1"], gn["p2"], gn["R"], gn["t"], gn["kappa"], gn["tau"])
    c = np.sum((gn["T_true"] @ pgo.connection_laplacian(measn, n)) * gn["T_true"])
    assert 0.0 < c < 0.2 * m
