"""CPU test of the host logic of dpgo_b200.rbcd (DeviceTeam / DeviceAgent): agent specs and neighbour
slots, the same-rank exchange, PGOAgent::iterate sequencing with Nesterov acceleration and the periodic
restart (ref: src/PGOAgent.cpp:376-432, 880-995), and the stream-ordered mode's bookkeeping -- with the
device behind the C-ABI replaced by a numpy stand-in that implements the same calls through the oracle.
The cost trace must equal the oracle Team's on the same schedule.  (That the real device calls compute
the same things is what the GPU tests check; this one pins the sequencing around them without a GPU.)"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import pgo, rbcd as orbcd


def _view(ptr, count, dtype):
    """numpy view of `count` items at a raw (CPU tensor) address, like a kernel would see it."""
    ct = {np.float64: C.c_double, np.int32: C.c_int32}[dtype]
    return np.ctypeslib.as_array((ct * count).from_address(ptr))


class _NumpyProblem:
    """Stand-in for dpgo_b200.api.DeviceProblem: same methods, numpy + the oracle's formulas."""

    ME = 0                       # own robot id inside the stand-in's LocalGraph; neighbour slot s = robot 1000 + s
    calls = []                   # (agent n, method) log shared by all instances of a test

    def __init__(self, n, d, r, device=0, stream=None):
        self.n, self.d, self.r = n, d, r
        self.tile = r * (d + 1)
        self.slots = {s: np.zeros((r, (d + 1) * n)) for s in range(4)}
        self.launches = 0
        self.pending = None
        self.closed = False

    # -- graph
    def set_private_edges(self, p1, p2, R, t, kappa, tau, weight=None):
        self.private = pgo.make_measurements(self.d, p1, p2, R, t, kappa, tau)

    def set_shared_edges(self, my_idx, nbr_slot, outgoing, R, t, kappa, tau, num_nbr_slots=0, weight=None):
        m = len(my_idx)
        out = np.asarray(outgoing).astype(bool)
        other_r = 1000 + np.asarray(nbr_slot, dtype=np.int64)
        me = np.full(m, self.ME, dtype=np.int64)
        zero = np.zeros(m, dtype=np.int64)
        my = np.asarray(my_idx, dtype=np.int64)
        self.shared = pgo.make_measurements(self.d, np.where(out, my, zero), np.where(out, zero, my), R, t, kappa, tau,
                                            r1=np.where(out, me, other_r), r2=np.where(out, other_r, me))
        self.num_nbr_slots = num_nbr_slots

    def finalize(self, build_precon=True):
        self.graph = pgo.LocalGraph(self.ME, self.d, self.r, self.n, self.private,
                                    self.shared if len(self.shared) else None)
        self.Q = pgo.construct_Q(self.graph)
        self.G = np.zeros((self.r, (self.d + 1) * self.n))
        self._lu = None

    # -- slots
    def slot_set(self, slot, X):
        self.slots[slot] = np.array(X, dtype=np.float64)

    def slot_get(self, slot):
        return self.slots[slot].copy()

    def slot_copy(self, dst, src):
        self.slots[dst] = self.slots[src].copy()

    def nesterov_update_Y(self, alpha):
        from dpgo_b200.api import SLOT_V, SLOT_X, SLOT_Y
        self.slots[SLOT_Y] = pgo.manifold_project((1 - alpha) * self.slots[SLOT_X] + alpha * self.slots[SLOT_V], self.d)

    def nesterov_update_V(self, gamma):
        from dpgo_b200.api import SLOT_V, SLOT_X, SLOT_Y
        self.slots[SLOT_V] = pgo.manifold_project(
            self.slots[SLOT_V] + gamma * (self.slots[SLOT_X] - self.slots[SLOT_Y]), self.d)

    # -- exchange
    def gather_tiles_dev(self, slot, num, idx_ptr, out_ptr):
        idx = _view(idx_ptr, num, np.int32)
        out = _view(out_ptr, num * self.tile, np.float64).reshape(num, self.tile)
        X = self.slots[slot]
        dh = self.d + 1
        for k, i in enumerate(idx):
            out[k] = X[:, i * dh:(i + 1) * dh].T.reshape(-1)      # column-major r x (d+1) tile

    def set_neighbor_poses_dev(self, ptr):
        dh = self.d + 1
        tiles = _view(ptr, max(self.num_nbr_slots, 1) * self.tile, np.float64)
        nbr = {(1000 + s, 0): tiles[s * self.tile:(s + 1) * self.tile].reshape(dh, self.r).T.copy()
               for s in range(self.num_nbr_slots)}
        self.G = pgo.construct_G(self.graph, nbr)

    # -- solve
    def _solve(self, src):
        from dpgo_b200.api import SLOT_X
        prob = pgo.QuadraticProblem(self.Q, self.G, self.d)
        if self._lu is not None:
            prob._lu = self._lu
        X, res = pgo.optimize(prob, self.slots[src])
        self._lu = prob._lu
        self.slots[SLOT_X] = X
        self.launches += 1
        return {"f_opt": res.fOpt, "inner_iters": res.inner_total, "outer_iters": res.outer}

    def optimize_slot(self, src, params=None):
        _NumpyProblem.calls.append((self.n, "optimize_slot"))
        return self._solve(src)

    def optimize_slot_async(self, src, params=None):
        _NumpyProblem.calls.append((self.n, "optimize_slot_async"))
        self.pending = self._solve(src)

    def optimize_result(self):
        _NumpyProblem.calls.append((self.n, "optimize_result"))
        assert self.pending is not None
        out, self.pending = self.pending, None
        return out

    def launch_count(self):
        return self.launches

    def close(self):
        self.closed = True


@pytest.fixture
def cpu_team(monkeypatch):
    from dpgo_b200 import rbcd
    real_device = torch.device
    monkeypatch.setattr(rbcd, "DeviceProblem", _NumpyProblem)
    monkeypatch.setattr(rbcd, "default_params", lambda **kw: None)
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu"))    # DeviceAgent's buffers on the host
    _NumpyProblem.calls = []
    return rbcd


@pytest.mark.parametrize("acceleration,rounds", [(True, 32), (False, 6)])
def test_colored_rounds_follow_the_oracle(datasets, cpu_team, acceleration, rounds):
    """32 accelerated rounds cross the periodic restart (iteration 29, restart_interval 30)."""
    meas, n, z = datasets("smallGrid3D")
    d, r, A = meas.d, 3, 5
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    ot = orbcd.Team(meas, n, A, r, acceleration=acceleration)
    ot.set_X(X0)
    colors = orbcd.robot_graph_coloring(ot.agents)
    traces = {}
    for mode in (False, True):
        team = cpu_team.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                                   stream=1, acceleration=acceleration)
        assert team.colors == colors
        team.set_async(mode)
        team.set_X(X0)
        trace = []
        for k in range(rounds):
            active = team.step_colored()
            assert active == colors[k % len(colors)]
            trace.append(2 * ot.central.f(team.assemble()))
        traces[mode] = trace
        results = {a: ag.result() for a, ag in team.agents.items()}
        assert all(res is not None and "f_opt" in res for res in results.values())
        assert all(ag.result() is results[a] for a, ag in team.agents.items())     # fetched once, then cached
        team.close()
        assert all(ag.prob.closed for ag in team.agents.values())
    ref = [ot.step_colored(colors, k)["cost"] for k in range(rounds)]
    assert np.allclose(traces[False], ref, rtol=1e-9, atol=0)
    assert traces[True] == traces[False]                       # same calls, same order: identical numbers
    names = {m for _, m in _NumpyProblem.calls}
    assert names == {"optimize_slot", "optimize_slot_async", "optimize_result"}
    # stream-ordered mode never asks for a result inside a round: one fetch per agent, at the end
    assert sum(m == "optimize_result" for _, m in _NumpyProblem.calls) == A


def test_all_agents_rounds_follow_the_oracle(datasets, cpu_team):
    meas, n, z = datasets("smallGrid3D")
    d, r, A = meas.d, 3, 5
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    ot = orbcd.Team(meas, n, A, r, acceleration=False)
    ot.set_X(X0)
    team = cpu_team.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                               stream=1, acceleration=False)
    team.set_async(True)
    team.set_X(X0)
    for _ in range(4):
        assert team.step_all() == list(range(A))
        want = ot.step_all()["cost"]
        got = 2 * ot.central.f(team.assemble())
        assert abs(got - want) <= 1e-9 * want
    with pytest.raises(ValueError):
        cpu_team.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                            stream=1, acceleration=True).step_all()
