"""GPU parity of the multi-agent path (BASELINE config 1: smallGrid3D, 5 agents, r = 5,
accelerated synchronous RBCD): device-resident agents vs the oracle's agents, same schedule."""
import numpy as np
import pytest

from oracle import pgo, rbcd as orbcd

pytestmark = pytest.mark.gpu


def _teams(datasets, name, A, r, acceleration=True, native=False):
    from dpgo_b200 import rbcd
    meas, n, z = datasets(name)
    d = meas.d
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    ot = orbcd.Team(meas, n, A, r, acceleration=acceleration)
    ot.set_X(X0)
    gt = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                         acceleration=acceleration, native_exchange=native)
    gt.set_X(X0)
    return meas, n, d, ot, gt


def test_G_from_neighbor_poses(datasets):
    """PoseGraph::constructG on device (cross block-CSR over the packed neighbour buffer)."""
    meas, n, d, ot, gt = _teams(datasets, "smallGrid3D", 5, 5)
    a = 2
    oa = ot.agents[a]
    for other in ot.agents:
        if other.id != a:
            oa.update_neighbor_poses(other.shared_pose_dict(False), False)
    G_ref = pgo.construct_G(oa.graph, oa.nbr)
    gt.exchange([a])
    ga = gt.agents[a]
    ga.prob.set_neighbor_poses_dev(ga.nbr.data_ptr())
    G = ga.prob.get_G()
    assert np.linalg.norm(G - G_ref) <= 1e-12 * np.linalg.norm(G_ref)
    # Q of the agent (private Laplacian + shared-edge diagonal terms) equals the oracle's
    import scipy.sparse as sp
    rowptr, colidx, blocks = ga.prob.get_Q_bsr()
    Qg = sp.bsr_matrix((blocks, colidx, rowptr), shape=oa.Q.shape).tocsr()
    assert abs(Qg - oa.Q).max() <= 1e-12 * abs(oa.Q).max()
    gt.close()


@pytest.mark.parametrize("acceleration", [True, False])
def test_colored_schedule_parity(datasets, acceleration):
    meas, n, d, ot, gt = _teams(datasets, "smallGrid3D", 5, 5, acceleration)
    colors = orbcd.robot_graph_coloring(ot.agents)
    assert colors == gt.colors
    central = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((5, 4 * n)), d)
    for k in range(35):     # crosses the periodic Nesterov restart at iteration 29
        so = ot.step_colored(colors, k)
        gt.step_colored()
        Xg = gt.assemble()
        cost_g = 2 * central.f(Xg)
        assert abs(cost_g - so["cost"]) <= 1e-8 * so["cost"], (k, cost_g, so["cost"])
    Xo = ot.assemble()
    assert np.linalg.norm(Xg - Xo) <= 1e-6 * np.linalg.norm(Xo)
    gt.close()


def test_greedy_schedule_parity(datasets):
    """The reference driver's schedule (one agent per iteration, greedy selection,
    examples/MultiRobotExample.cpp:170-247)."""
    import dpgo_b200
    meas, n, d, ot, gt = _teams(datasets, "smallGrid3D", 5, 5)
    central_g = dpgo_b200.problem_from_measurements(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau,
                                                    n, d, 5, build_precon=False)
    sel = 0
    for k in range(20):
        so = ot.step_greedy()
        assert so["robot"] == sel
        gt.step_single(sel)
        Xg = gt.assemble()
        sel, gn = gt.greedy_select(central_g, Xg)
        assert abs(2 * central_g.f(Xg) - so["cost"]) <= 1e-8 * so["cost"]
        assert abs(gn - so["gradnorm"]) <= 1e-6 * so["gradnorm"]
    gt.close(); central_g.close()


def test_measurement_errors_parity(datasets):
    """computeMeasurementError (src/DPGO_utils.cpp:501-507) of every private / shared edge by the
    device edge kernel vs the oracle's per-edge loop, at a non-trivial iterate."""
    meas, n, d, ot, gt = _teams(datasets, "smallGrid3D", 5, 5, acceleration=False)
    colors = orbcd.robot_graph_coloring(ot.agents)
    for k in range(2):
        ot.step_colored(colors, k)
        gt.step_colored()
    for a in ot.agents:
        for other in ot.agents:
            if other.id != a.id:
                a.update_neighbor_poses(other.shared_pose_dict(False), False)
    gt.exchange(list(range(5)))
    for a in ot.agents:
        rp, rs = a.measurement_residuals()
        ga = gt.agents[a.id]
        ep, es = ga.prob.measurement_errors(0, ga.nbr.data_ptr())
        assert np.allclose(ep, rp ** 2, rtol=1e-9, atol=1e-12)
        assert np.allclose(es, rs ** 2, rtol=1e-9, atol=1e-12)
        assert np.array_equal(ga.spec.priv_fixed, a.graph.private.fixed)
        assert np.array_equal(ga.spec.shared_fixed, a.graph.shared.fixed)
    gt.close()


@pytest.mark.parametrize("native", [False, True])
def test_gnc_weight_update_parity(datasets, native):
    """native = the public poses move through dpgo_exchange (the path bench.py times), else through the Python exchange.
    BASELINE config 5: city10000 (2D), 4 agents, r = 3, GNC_TLS loop-closure weights
    (PGOAgent::updateMeasurementWeights, src/PGOAgent.cpp:1104-1142; RobustCost defaults): the
    device path (edge-error kernel, Q and two-level preconditioner rebuilt on the device after
    every weight update) follows the oracle's agents through two weight updates."""
    meas, n, d, ot, gt = _teams(datasets, "city10000", 4, 3, acceleration=False, native=native)
    colors = orbcd.robot_graph_coloring(ot.agents)
    assert colors == gt.colors
    central = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((3, 3 * n)), d)
    k = 0
    for stage, rounds in enumerate((4, 4, 2)):
        for _ in range(rounds):
            so = ot.step_colored(colors, k)
            gt.step_colored()
            k += 1
            cost_g = 2 * central.f(gt.assemble())
            assert abs(cost_g - so["cost"]) <= 1e-7 * so["cost"], (stage, k, cost_g, so["cost"])
        if stage < 2:
            wo = ot.update_weights()
            wg = gt.update_weights()
            for a in range(4):
                assert np.allclose(wg[a][0], wo[a][0], rtol=0, atol=1e-6), (stage, a)
                assert np.allclose(wg[a][1], wo[a][1], rtol=0, atol=1e-6), (stage, a)
                lc = ~ot.agents[a].graph.private.fixed
                assert np.all(wo[a][0][~lc] == 1.0) and wo[a][0][lc].min() < 0.5   # weights really moved
            assert abs(gt.robust[0].mu - ot.robust[0].mu) < 1e-15
    Xo, Xg = ot.assemble(), gt.assemble()
    assert np.linalg.norm(Xg - Xo) <= 1e-6 * np.linalg.norm(Xo)
    gt.close()


@pytest.mark.parametrize("native", [False, True])
def test_all_agents_schedule_parity(datasets, native):
    """BASELINE config 4: torus3D, 8 agents, r = 5, every agent optimizes in every round with the
    poses of the previous round (the equal-rate instance of asynchronous parallel RBCD,
    src/PGOAgent.cpp:486-499; no acceleration)."""
    meas, n, d, ot, gt = _teams(datasets, "torus3D", 8, 5, acceleration=False, native=native)
    central = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((5, 4 * n)), d)
    prev = None
    for k in range(6):
        so = ot.step_all()
        gt.step_all()
        Xg = gt.assemble()
        cost_g = 2 * central.f(Xg)
        assert abs(cost_g - so["cost"]) <= 1e-8 * so["cost"], (k, cost_g, so["cost"])
        prev = so["cost"]
    Xo = ot.assemble()
    assert np.linalg.norm(Xg - Xo) <= 1e-6 * np.linalg.norm(Xo)
    gt.close()


@pytest.mark.parametrize("name,r", [("smallGrid3D", 5), ("city10000", 3)])
def test_device_weight_refresh_bitwise(datasets, name, r):
    """dpgo_update_weights with an unchanged pattern re-weights Q and the cross blocks on the device (k_refresh_q /
    k_refresh_cross restate the host assembly entry by entry, same operations in the same order, no contraction):
    the refreshed Q, the linear term built from neighbour poses and the preconditioner equal those of a handle whose
    matrices were assembled on the host from the same weights -- bit for bit.
    ref: PoseGraph::clearDataMatrices after weight updates, src/PGOAgent.cpp:1062-1142."""
    import dpgo_b200
    meas, n, _ = datasets(name)
    d = meas.d
    rng = np.random.default_rng(5)
    m = len(meas.p1)
    n_a = n // 2                                   # agent = first half of the poses; edges across the cut are shared
    priv = (meas.p1 < n_a) & (meas.p2 < n_a)
    cut = (meas.p1 < n_a) != (meas.p2 < n_a)
    out = meas.p1[cut] < n_a
    my = np.where(out, meas.p1[cut], meas.p2[cut]).astype(np.int32)
    slots = np.arange(int(cut.sum()), dtype=np.int32)
    w1p, w1s = rng.uniform(0.1, 1.0, int(priv.sum())), rng.uniform(0.1, 1.0, int(cut.sum()))
    w2p, w2s = rng.uniform(0.0, 1.0, int(priv.sum())), rng.uniform(0.0, 1.0, int(cut.sum()))
    w2p[::7] = 0.0                                 # rejected measurements

    def build(wp, ws):
        gp = dpgo_b200.DeviceProblem(n_a, d, r)
        gp.set_private_edges(meas.p1[priv], meas.p2[priv], meas.R[priv], meas.t[priv], meas.kappa[priv], meas.tau[priv], wp)
        gp.set_shared_edges(my, slots, out.astype(np.uint8), meas.R[cut], meas.t[cut], meas.kappa[cut], meas.tau[cut],
                            len(slots), ws)
        gp.finalize(True)
        return gp

    nbr = rng.standard_normal((len(slots), r, d + 1))
    V = np.asfortranarray(rng.standard_normal((r, (d + 1) * n_a)))
    X = pgo.manifold_project(rng.standard_normal((r, (d + 1) * n_a)), d)
    a = build(w1p, w1s)
    a.update_weights(w2p, w2s, True)               # device refresh
    b = build(w2p, w2s)                            # host assembly
    qa, qb = a.get_Q_bsr(), b.get_Q_bsr()
    assert all(np.array_equal(x, y) for x, y in zip(qa, qb))
    a.set_neighbor_poses(nbr); b.set_neighbor_poses(nbr)
    assert np.array_equal(a.get_G(), b.get_G())
    assert np.array_equal(a.precon(X, V), b.precon(X, V))
    a.update_weights(w1p, w1s, True)               # and back: equals the first assembly
    c = build(w1p, w1s)
    assert all(np.array_equal(x, y) for x, y in zip(a.get_Q_bsr(), c.get_Q_bsr()))
    for gp in (a, b, c):
        gp.close()
