"""GPU parity of the multi-agent path (BASELINE config 1: smallGrid3D, 5 agents, r = 5,
accelerated synchronous RBCD): device-resident agents vs the oracle's agents, same schedule."""
import numpy as np
import pytest

from oracle import pgo, rbcd as orbcd

pytestmark = pytest.mark.gpu


def _teams(datasets, name, A, r, acceleration=True):
    from dpgo_b200 import rbcd
    meas, n, z = datasets(name)
    d = meas.d
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    ot = orbcd.Team(meas, n, A, r, acceleration=acceleration)
    ot.set_X(X0)
    gt = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                         acceleration=acceleration)
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
