"""numpy restatement of the multi-agent RBCD loop (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows PGOAgent::iterate / updateX / Nesterov updates (src/PGOAgent.cpp:376-432, 880-995) and
the driver's partition / exchange / selection logic (examples/MultiRobotExample.cpp:71-255).
"""
from __future__ import annotations

import math

import numpy as np

from . import pgo


def partition(meas: pgo.Measurements, n: int, num_robots: int):
    """examples/MultiRobotExample.cpp:71-119: contiguous equal split, last robot takes the
    remainder.  Returns (ranges, list of (private, shared) Measurements per robot).  Every
    returned Measurements carries `fixed` = the reader's fixedWeight flag (consecutive GLOBAL
    pose ids, src/DPGO_utils.cpp:178,232), which the driver copies along."""
    per = n // num_robots
    assert per > 0
    starts = [k * per for k in range(num_robots)]
    ends = [(k + 1) * per for k in range(num_robots)]
    ends[-1] = n
    rob = np.minimum(np.arange(n) // per, num_robots - 1)
    loc = np.arange(n) - np.array(starts)[rob]
    r1, r2 = rob[meas.p1], rob[meas.p2]
    l1, l2 = loc[meas.p1], loc[meas.p2]
    out = []
    for k in range(num_robots):
        priv = np.where((r1 == k) & (r2 == k))[0]
        sh = np.where(((r1 == k) | (r2 == k)) & (r1 != r2))[0]
        P = meas.subset(priv)
        P = pgo.Measurements(P.d, r1[priv], l1[priv], r2[priv], l2[priv], P.R, P.t, P.kappa, P.tau, P.weight)
        S = meas.subset(sh)
        S = pgo.Measurements(S.d, r1[sh], l1[sh], r2[sh], l2[sh], S.R, S.t, S.kappa, S.tau, S.weight)
        P.fixed = (meas.p1[priv] + 1 == meas.p2[priv])
        S.fixed = (meas.p1[sh] + 1 == meas.p2[sh])
        out.append((P, S))
    return list(zip(starts, ends)), out


class Agent:
    """The subset of PGOAgent that the synchronous RBCD loop exercises."""

    def __init__(self, rid, d, r, n, private, shared, num_robots, acceleration=True,
                 restart_interval=30, params: pgo.ROptParameters | None = None):
        self.id, self.d, self.r, self.n = rid, d, r, n
        self.graph = pgo.LocalGraph(rid, d, r, n, private, shared)
        self.Q = pgo.construct_Q(self.graph)
        self.num_robots = num_robots
        self.acceleration = acceleration
        self.restart_interval = restart_interval
        self.params = params or pgo.ROptParameters()
        self.iteration = 0
        self.nbr, self.nbr_aux = {}, {}
        self.X = None
        self.last_result = None
        self._prob = None

    # -- state
    def set_X(self, X):                         # PGOAgent::setX :52-63
        self.X = X.copy()
        if self.acceleration:                   # initializeAcceleration :899-908
            self.XPrev = self.X.copy()
            self.gamma = 0.0
            self.alpha = 0.0
            self.V = self.X.copy()
            self.Y = self.X.copy()

    def _pose(self, M, i):
        dh = self.d + 1
        return M[:, i * dh:(i + 1) * dh].copy()

    def shared_pose_dict(self, aux=False):      # getSharedPoseDict :97-110 / aux :132-146
        M = self.Y if aux else self.X
        return {(self.id, i): self._pose(M, i) for i in self.graph.public_pose_ids()}

    def update_neighbor_poses(self, poses, aux=False):   # :650-702
        need = set(self.graph.neighbor_pose_ids())
        tgt = self.nbr_aux if aux else self.nbr
        for k, v in poses.items():
            if k in need:
                tgt[k] = v

    # -- robust weights
    def measurement_residuals(self):
        """sqrt(computeMeasurementError) of every private / shared edge at X and the neighbours' X
        (PGOAgent::computeMeasurementResidual, src/PGOAgent.cpp:1062-1102)."""
        d = self.d
        rot = lambda T: T[:, :d]
        tr = lambda T: T[:, d]
        P, S = self.graph.private, self.graph.shared
        rp = np.zeros(len(P))
        for k in range(len(P)):
            T1, T2 = self._pose(self.X, int(P.p1[k])), self._pose(self.X, int(P.p2[k]))
            rp[k] = math.sqrt(pgo.measurement_error(P.R[k], P.t[k], P.kappa[k], P.tau[k], rot(T1), tr(T1), rot(T2), tr(T2)))
        rs = np.zeros(0 if S is None else len(S))
        for k in range(len(rs)):
            if S.r1[k] == self.id:
                T1, T2 = self._pose(self.X, int(S.p1[k])), self.nbr[(int(S.r2[k]), int(S.p2[k]))]
            else:
                T1, T2 = self.nbr[(int(S.r1[k]), int(S.p1[k]))], self._pose(self.X, int(S.p2[k]))
            rs[k] = math.sqrt(pgo.measurement_error(S.R[k], S.t[k], S.kappa[k], S.tau[k], rot(T1), tr(T1), rot(T2), tr(T2)))
        return rp, rs

    def update_measurement_weights(self, robust: pgo.RobustCost):
        """PGOAgent::updateMeasurementWeights, src/PGOAgent.cpp:1104-1142 (robustOptNumResets = 0:
        the trajectory estimate is kept): re-weight every loop closure whose weight is not
        fixed, advance the GNC schedule, drop the data matrices, restart the acceleration."""
        rp, rs = self.measurement_residuals()
        P, S = self.graph.private, self.graph.shared
        for k in range(len(P)):
            if not P.fixed[k]:
                P.weight[k] = robust.weight(rp[k])
        for k in range(len(rs)):
            if not S.fixed[k]:
                S.weight[k] = robust.weight(rs[k])
        robust.update()
        self.Q = pgo.construct_Q(self.graph)            # clearDataMatrices
        self._prob = None
        self._cpu = None
        if self.acceleration:                           # initializeAcceleration :899-908
            self.XPrev = self.X.copy()
            self.gamma = self.alpha = 0.0
            self.V = self.X.copy()
            self.Y = self.X.copy()

    # -- iterate
    def _update_X(self, do_opt, acceleration):  # updateX :938-995
        if not do_opt:
            if acceleration:
                self.X = self.Y.copy()
            return True
        G = pgo.construct_G(self.graph, self.nbr_aux if acceleration else self.nbr)
        X0 = self.Y if acceleration else self.X
        if getattr(self, "use_cpu_port", False):
            # compiled single-thread restatement (oracle/cpu_port): same algorithm, C++ speed
            from .cpu_port import CpuProblem
            if getattr(self, "_cpu", None) is None:
                self._cpu = CpuProblem(self.Q, G, self.d)
            self._cpu.set_G(G)
            p = self.params
            self.X, res = self._cpu.optimize(X0, p.gradnorm_tol, p.RTR_iterations, p.RTR_tCG_iterations,
                                             p.RTR_initial_radius)
            self.last_result = res
            return True
        prob = pgo.QuadraticProblem(self.Q, G, self.d)
        if self._prob is not None:
            prob._lu = self._prob._lu           # Q unchanged => preconditioner unchanged
        X0 = self.Y if acceleration else self.X
        self.X, self.last_result = pgo.optimize(prob, X0, self.params)
        self._prob = prob
        return True

    def iterate(self, do_opt=True):             # :376-432
        self.iteration += 1
        self.XPrev = self.X.copy()
        if self.acceleration:
            R = self.num_robots
            self.gamma = (1 + math.sqrt(1 + 4 * R ** 2 * self.gamma ** 2)) / (2 * R)   # :910-914
            self.alpha = 1 / (self.gamma * R)                                          # :916-920
            self.Y = pgo.manifold_project((1 - self.alpha) * self.X + self.alpha * self.V, self.d)
            ok = self._update_X(do_opt, True)
            self.V = pgo.manifold_project(self.V + self.gamma * (self.X - self.Y), self.d)
            if (self.iteration + 1) % self.restart_interval == 0:                      # :880-897
                self.X = self.XPrev.copy()
                self._update_X(do_opt, False)
                self.V = self.X.copy()
                self.Y = self.X.copy()
                self.gamma = 0.0
                self.alpha = 0.0
            return ok
        return self._update_X(do_opt, False)


def robot_graph_coloring(agents):
    """Greedy colouring of the robot graph: agents of one colour share no edge, so
    updating them concurrently equals updating them one after another."""
    color = {}
    for a in agents:
        used = {color[n] for n in a.graph.neighbor_ids() if n in color}
        c = 0
        while c in used:
            c += 1
        color[a.id] = c
    ncol = max(color.values()) + 1 if color else 1
    return [[a.id for a in agents if color[a.id] == c] for c in range(ncol)]


class Team:
    def __init__(self, meas, n, num_robots, r, acceleration=True, params=None, restart_interval=30):
        self.d, self.r, self.n, self.R = meas.d, r, n, num_robots
        self.meas = meas
        self.ranges, parts = partition(meas, n, num_robots)
        self.agents = [Agent(k, self.d, r, e - s, parts[k][0], parts[k][1], num_robots,
                             acceleration, restart_interval, params)
                       for k, (s, e) in enumerate(self.ranges)]
        self.acceleration = acceleration
        Qc = pgo.connection_laplacian(meas, n)
        self.central = pgo.QuadraticProblem(Qc, np.zeros((r, (self.d + 1) * n)), self.d)
        self.selected = 0

    def set_X(self, X):
        dh = self.d + 1
        for a, (s, e) in zip(self.agents, self.ranges):
            a.set_X(X[:, s * dh:e * dh])

    def assemble(self):
        return np.concatenate([a.X for a in self.agents], axis=1)

    def _exchange_to(self, sel):
        for other in self.agents:
            if other.id == sel.id:
                continue
            sel.update_neighbor_poses(other.shared_pose_dict(False), False)
            if self.acceleration:
                sel.update_neighbor_poses(other.shared_pose_dict(True), True)

    def step_greedy(self):
        """One iteration of examples/MultiRobotExample.cpp:170-247 (one agent optimizes)."""
        sel = self.agents[self.selected]
        for a in self.agents:
            if a.id != sel.id:
                a.iterate(False)
        self._exchange_to(sel)
        sel.iterate(True)
        X = self.assemble()
        RG = self.central.rgrad(X)
        cost = 2 * self.central.f(X)
        gn = float(np.linalg.norm(RG))
        who = sel.id
        if sel.graph.neighbor_ids():
            dh = self.d + 1
            norms = [np.linalg.norm(RG[:, s * dh:e * dh]) for (s, e) in self.ranges]
            self.selected = int(np.argmax(norms))
        return dict(robot=who, cost=cost, gradnorm=gn)

    def step_colored(self, colors, k):
        """Parallel block schedule: all agents of colour k % ncol optimize."""
        active = set(colors[k % len(colors)])
        for a in self.agents:
            if a.id not in active:
                a.iterate(False)
        for a in self.agents:
            if a.id in active:
                self._exchange_to(a)
        for a in self.agents:
            if a.id in active:
                a.iterate(True)
        X = self.assemble()
        return dict(robots=sorted(active), cost=2 * self.central.f(X),
                    gradnorm=self.central.rgrad_norm(X))

    def step_all(self):
        """Every agent optimizes in the same round from the poses its neighbours published at the
        end of the previous one -- the deterministic (equal-rate, no jitter) instance of the
        asynchronous parallel mode (src/PGOAgent.cpp:486-499; acceleration is not allowed there,
        :477)."""
        assert not self.acceleration
        for a in self.agents:
            self._exchange_to(a)
        for a in self.agents:
            a.iterate(True)
        X = self.assemble()
        return dict(cost=2 * self.central.f(X), gradnorm=self.central.rgrad_norm(X))

    def update_weights(self):
        """All agents refresh their neighbours' poses and re-weight their loop closures (each
        agent owns a RobustCost with the same schedule, as every PGOAgent does)."""
        if not hasattr(self, "robust"):
            self.robust = [pgo.RobustCost() for _ in self.agents]
        for a in self.agents:
            for other in self.agents:
                if other.id != a.id:
                    a.update_neighbor_poses(other.shared_pose_dict(False), False)
        for a, rc in zip(self.agents, self.robust):
            a.update_measurement_weights(rc)
        # the centralized cost the driver reports uses the current weights of the global graph
        return [(a.graph.private.weight.copy(), a.graph.shared.weight.copy()) for a in self.agents]
