"""numpy/scipy restatement of the reference's local-solve hot path (TEST INFRASTRUCTURE).

All citations are file:line under /root/reference.  See oracle/__init__.py for the parity
status ("parity unpinned" for ROPTLIB-internal control flow).

Conventions (same as the reference, tests/testEigenMap.cpp:12-36 pins the layout):
  X is r x (d+1)n, pose i occupies columns (d+1)i .. (d+1)i+d; the first d columns are the
  Stiefel block Y_i in St(d, r), the last column is the translation p_i.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------------------
# measurements
# --------------------------------------------------------------------------------------
@dataclass
class Measurements:
    """Struct-of-arrays version of std::vector<RelativeSEMeasurement>
    (include/DPGO/RelativeSEMeasurement.h:23-88)."""
    d: int
    r1: np.ndarray      # robot id of tail   (int64, m)
    p1: np.ndarray      # frame id of tail
    r2: np.ndarray
    p2: np.ndarray
    R: np.ndarray       # (m, d, d)
    t: np.ndarray       # (m, d)
    kappa: np.ndarray   # (m,)
    tau: np.ndarray     # (m,)
    weight: np.ndarray  # (m,)

    def __len__(self):
        return len(self.p1)

    def subset(self, idx):
        idx = np.asarray(idx)
        return Measurements(self.d, self.r1[idx], self.p1[idx], self.r2[idx], self.p2[idx],
                            self.R[idx], self.t[idx], self.kappa[idx], self.tau[idx],
                            self.weight[idx])

    @staticmethod
    def concat(parts):
        parts = [p for p in parts if p is not None]
        d = parts[0].d
        cat = lambda a: np.concatenate([getattr(p, a) for p in parts], axis=0)
        return Measurements(d, cat("r1"), cat("p1"), cat("r2"), cat("p2"), cat("R"), cat("t"),
                            cat("kappa"), cat("tau"), cat("weight"))


def make_measurements(d, p1, p2, R, t, kappa, tau, r1=None, r2=None, weight=None):
    m = len(p1)
    z = np.zeros(m, dtype=np.int64)
    return Measurements(d,
                        z.copy() if r1 is None else np.asarray(r1, dtype=np.int64),
                        np.asarray(p1, dtype=np.int64),
                        z.copy() if r2 is None else np.asarray(r2, dtype=np.int64),
                        np.asarray(p2, dtype=np.int64),
                        np.asarray(R, dtype=np.float64).reshape(m, d, d),
                        np.asarray(t, dtype=np.float64).reshape(m, d),
                        np.asarray(kappa, dtype=np.float64),
                        np.asarray(tau, dtype=np.float64),
                        np.ones(m) if weight is None else np.asarray(weight, dtype=np.float64))


def quat_to_rot(qx, qy, qz, qw):
    """Eigen::Quaterniond(w,x,y,z).toRotationMatrix() -- src/DPGO_utils.cpp:214.
    Eigen does NOT normalize; it uses the standard 2xy/2wz formula on the raw coefficients."""
    tx, ty, tz = 2 * qx, 2 * qy, 2 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def read_g2o(path):
    """src/DPGO_utils.cpp:113-257.  Returns (Measurements, num_poses)."""
    p1, p2, Rs, ts, kap, tau = [], [], [], [], [], []
    d = None
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "EDGE_SE2":
                v = [float(x) for x in tok[3:]]
                i, j = int(tok[1]), int(tok[2])
                dx, dy, dth, I11, I12, I13, I22, I23, I33 = v[:9]
                d = 2
                c, s = math.cos(dth), math.sin(dth)
                Rs.append(np.array([[c, -s], [s, c]]))
                ts.append(np.array([dx, dy]))
                cov = np.array([[I11, I12], [I12, I22]])
                tau.append(2.0 / np.trace(np.linalg.inv(cov)))     # :174
                kap.append(I33)                                     # :176
            elif tok[0] == "EDGE_SE3:QUAT":
                v = [float(x) for x in tok[3:]]
                i, j = int(tok[1]), int(tok[2])
                dx, dy, dz, qx, qy, qz, qw = v[:7]
                (I11, I12, I13, I14, I15, I16, I22, I23, I24, I25, I26, I33, I34, I35, I36,
                 I44, I45, I46, I55, I56, I66) = v[7:28]
                d = 3
                Rs.append(quat_to_rot(qx, qy, qz, qw))
                ts.append(np.array([dx, dy, dz]))
                tc = np.array([[I11, I12, I13], [I12, I22, I23], [I13, I23, I33]])
                tau.append(3.0 / np.trace(np.linalg.inv(tc)))       # :223
                rc = np.array([[I44, I45, I46], [I45, I55, I56], [I46, I56, I66]])
                kap.append(3.0 / (2.0 * np.trace(np.linalg.inv(rc))))  # :230
            elif tok[0] in ("VERTEX_SE2", "VERTEX_SE3:QUAT"):
                continue
            else:
                raise ValueError("unrecognized g2o token " + tok[0])
            p1.append(i)
            p2.append(j)
    meas = make_measurements(d, p1, p2, np.array(Rs), np.array(ts), kap, tau)
    n = int(max(meas.p1.max(), meas.p2.max())) + 1
    return meas, n


# --------------------------------------------------------------------------------------
# data matrices
# --------------------------------------------------------------------------------------
def _homog(R, t):
    d = R.shape[0]
    T = np.zeros((d + 1, d + 1))
    T[:d, :d] = R
    T[:d, d] = t
    T[d, d] = 1.0
    return T


def _omega(d, w, kappa, tau):
    return np.diag([w * kappa] * d + [w * tau])


def connection_laplacian(meas: Measurements, n: int) -> sp.csr_matrix:
    """constructConnectionLaplacianSE, src/DPGO_utils.cpp:272-344: Q = A Omega A^T with
    A(i,k) = -T_ij, A(j,k) = +I."""
    d = meas.d
    dh = d + 1
    rows, cols, vals = [], [], []
    ar = np.arange(dh)
    for k in range(len(meas)):
        i, j = int(meas.p1[k]), int(meas.p2[k])
        T = _homog(meas.R[k], meas.t[k])
        Om = _omega(d, meas.weight[k], meas.kappa[k], meas.tau[k])
        for (a, b, blk) in ((i, i, T @ Om @ T.T), (i, j, -T @ Om), (j, i, -Om @ T.T), (j, j, Om)):
            rr, cc = np.meshgrid(a * dh + ar, b * dh + ar, indexing="ij")
            rows.append(rr.ravel()); cols.append(cc.ravel()); vals.append(blk.ravel())
    if not rows:
        return sp.csr_matrix((dh * n, dh * n))
    Q = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(dh * n, dh * n)).tocsr()
    return Q


@dataclass
class LocalGraph:
    """One agent's PoseGraph (src/PoseGraph.cpp).  `private` holds odometry + private loop
    closures with local frame ids; `shared` keeps (r1,p1,r2,p2) global robot/frame ids."""
    robot_id: int
    d: int
    r: int
    n: int
    private: Measurements
    shared: Measurements | None = None
    priors: dict = field(default_factory=dict)   # idx -> r x (d+1) pose
    prior_kappa: float = 10000.0                 # src/PoseGraph.cpp:17-18
    prior_tau: float = 100.0

    def public_pose_ids(self):
        """myPublicPoseIDs (src/PoseGraph.cpp:615-630): sorted set of my frame ids in shared edges."""
        if self.shared is None or len(self.shared) == 0:
            return []
        s = self.shared
        ids = set(int(p) for p, rr in zip(s.p1, s.r1) if rr == self.robot_id)
        ids |= set(int(p) for p, rr in zip(s.p2, s.r2) if rr == self.robot_id)
        return sorted(ids)

    def neighbor_pose_ids(self):
        if self.shared is None or len(self.shared) == 0:
            return []
        s = self.shared
        ids = set((int(rr), int(p)) for p, rr in zip(s.p1, s.r1) if rr != self.robot_id)
        ids |= set((int(rr), int(p)) for p, rr in zip(s.p2, s.r2) if rr != self.robot_id)
        return sorted(ids)

    def neighbor_ids(self):
        return sorted(set(rid for rid, _ in self.neighbor_pose_ids()))


def construct_Q(g: LocalGraph) -> sp.csr_matrix:
    """PoseGraph::constructQ, src/PoseGraph.cpp:381-491."""
    d, dh, n = g.d, g.d + 1, g.n
    Q = connection_laplacian(g.private, n).tolil()
    if g.shared is not None:
        s = g.shared
        for k in range(len(s)):
            T = _homog(s.R[k], s.t[k])
            Om = _omega(d, s.weight[k], s.kappa[k], s.tau[k])
            if s.r1[k] == g.robot_id:      # outgoing edge :433-434
                i = int(s.p1[k])
                Q[i * dh:(i + 1) * dh, i * dh:(i + 1) * dh] += T @ Om @ T.T
            else:                           # incoming edge :457
                i = int(s.p2[k])
                Q[i * dh:(i + 1) * dh, i * dh:(i + 1) * dh] += Om
    for idx in g.priors:                    # :462-469
        Om = _omega(d, 1.0, g.prior_kappa, g.prior_tau)
        Q[idx * dh:(idx + 1) * dh, idx * dh:(idx + 1) * dh] += Om
    return Q.tocsr()


def construct_G(g: LocalGraph, neighbor_poses: dict) -> np.ndarray:
    """PoseGraph::constructG, src/PoseGraph.cpp:493-580.  neighbor_poses maps
    (robot_id, frame_id) -> r x (d+1) lifted pose."""
    d, dh, n, r = g.d, g.d + 1, g.n, g.r
    G = np.zeros((r, dh * n))
    if g.shared is not None:
        s = g.shared
        for k in range(len(s)):
            T = _homog(s.R[k], s.t[k])
            Om = _omega(d, s.weight[k], s.kappa[k], s.tau[k])
            if s.r1[k] == g.robot_id:      # outgoing :536-537
                Xj = neighbor_poses[(int(s.r2[k]), int(s.p2[k]))]
                i = int(s.p1[k])
                G[:, i * dh:(i + 1) * dh] += -Xj @ Om @ T.T
            else:                           # incoming :561-562
                Xi = neighbor_poses[(int(s.r1[k]), int(s.p1[k]))]
                i = int(s.p2[k])
                G[:, i * dh:(i + 1) * dh] += -Xi @ T @ Om
    for idx, P in g.priors.items():         # :573-574
        Om = _omega(d, 1.0, g.prior_kappa, g.prior_tau)
        G[:, idx * dh:(idx + 1) * dh] += -P @ Om
    return G


# --------------------------------------------------------------------------------------
# manifold (St(d,r) x R^r)^n   -- ROPTLIB semantics, restated (see oracle/__init__.py)
# --------------------------------------------------------------------------------------
def _tiles(X, d):
    r = X.shape[0]
    return X.T.reshape(-1, d + 1, r).transpose(0, 2, 1)      # shape (n, r, d+1)


def _untile(Tl):
    n, r, dh = Tl.shape
    return Tl.transpose(0, 2, 1).reshape(n * dh, r).T.copy()


def tangent_project(X, V, d):
    """ProductManifold::Projection with Stiefel extrinsic projection (Euclidean metric):
    xi_Y = V_Y - Y sym(Y^T V_Y); translation untouched.  SURVEY 8(a) C2."""
    Xt, Vt = _tiles(X, d), _tiles(V, d)
    Y, W = Xt[:, :, :d], Vt[:, :, :d]
    M = np.einsum("nra,nrb->nab", Y, W)
    S = 0.5 * (M + M.transpose(0, 2, 1))
    out = Vt.copy()
    out[:, :, :d] = W - np.einsum("nra,nab->nrb", Y, S)
    return _untile(out)


def retract_qf(X, Eta, d):
    """ProductManifold::Retraction with Stiefel QF retraction: Y+ = qf(Y + eta_Y) with the
    diagonal of R forced positive; p+ = p + eta_p.  SURVEY 8(a) C4."""
    Zt = _tiles(X + Eta, d)
    out = Zt.copy()
    Qm, Rm = np.linalg.qr(Zt[:, :, :d])                     # stacked thin QR, one per pose
    sg = np.sign(np.diagonal(Rm, axis1=1, axis2=2))
    sg[sg == 0] = 1.0
    out[:, :, :d] = Qm * sg[:, None, :]
    return _untile(out)


def project_stiefel(M):
    """projectToStiefelManifold, src/DPGO_utils.cpp:480-486: U V^T of the thin SVD."""
    U, _, Vt = np.linalg.svd(M, full_matrices=False)
    return U @ Vt


def project_rotation(M):
    """projectToRotationGroup, src/DPGO_utils.cpp:464-478."""
    U, _, Vt = np.linalg.svd(M)
    if np.linalg.det(U) * np.linalg.det(Vt) > 0:
        return U @ Vt
    U = U.copy()
    U[:, -1] *= -1
    return U @ Vt


def manifold_project(X, d):
    """LiftedSEManifold::project, src/manifold/LiftedSEManifold.cpp:34-45."""
    Xt = _tiles(X, d).copy()
    # project_stiefel of every pose at once (numpy's svd works on stacked matrices)
    U, _, Vt = np.linalg.svd(Xt[:, :, :d], full_matrices=False)
    Xt[:, :, :d] = U @ Vt
    return _untile(Xt)


# --------------------------------------------------------------------------------------
# QuadraticProblem
# --------------------------------------------------------------------------------------
class QuadraticProblem:
    """src/QuadraticProblem.cpp.  f(X) = 0.5 <Q, X^T X> + <X, G>."""

    def __init__(self, Q: sp.csr_matrix, G: np.ndarray, d: int, precon_shift: float = 0.1):
        self.Q = Q.tocsr()
        self.G = G
        self.d = d
        self.r = G.shape[0]
        self.n = G.shape[1] // (d + 1)
        self.shift = precon_shift
        self._lu = None
        self.counters = dict(qx=0, precon=0, proj=0, retract=0)

    # -- Q application (X Q, Q symmetric)
    def XQ(self, X):
        self.counters["qx"] += 1
        return (self.Q @ X.T).T

    def f(self, X):                               # :29-41
        return 0.5 * float(np.sum(self.XQ(X) * X)) + float(np.sum(X * self.G))

    def egrad(self, X):                           # :43-47
        return self.XQ(X) + self.G

    def ehess(self, V):                           # :49-54
        return self.XQ(V)

    def precon_solve(self, V):
        """(Q + 0.1 I)^{-1} applied to the columns of V^T -- src/PoseGraph.cpp:598-613."""
        if self._lu is None:
            P = (self.Q + self.shift * sp.identity(self.Q.shape[0])).tocsc()
            self._lu = spla.splu(P)
        self.counters["precon"] += 1
        return self._lu.solve(np.ascontiguousarray(V.T)).T

    def precondition(self, X, V):                 # :56-69
        return tangent_project(X, self.precon_solve(V), self.d)

    def rgrad(self, X):                           # :71-79
        return tangent_project(X, self.egrad(X), self.d)

    def rgrad_norm(self, X):                      # :81-83
        return float(np.linalg.norm(self.rgrad(X)))

    def rhess(self, X, EG, V):
        """Riemannian Hessian-vector (Stiefel::EucHvToHv, Euclidean metric): per pose
        Proj_Y( (VQ)_Y - V_Y sym(Y^T EG_Y) ), translation (VQ)_p.  SURVEY 8(a) C3."""
        d = self.d
        HV = self.ehess(V)
        Xt, Et, Vt, Ht = _tiles(X, d), _tiles(EG, d), _tiles(V, d), _tiles(HV, d).copy()
        M = np.einsum("nra,nrb->nab", Xt[:, :, :d], Et[:, :, :d])
        S = 0.5 * (M + M.transpose(0, 2, 1))
        Ht[:, :, :d] -= np.einsum("nra,nab->nrb", Vt[:, :, :d], S)
        return tangent_project(X, _untile(Ht), d)


# --------------------------------------------------------------------------------------
# QuadraticOptimizer  (RTR / tCG restated from ROPTLIB, see module docstring)
# --------------------------------------------------------------------------------------
TCG_LCON, TCG_SCON, TCG_NEGCURV, TCG_EXCREGION, TCG_MAXITER = 0, 1, 2, 3, 4


@dataclass
class ROptParameters:          # include/DPGO/DPGO_types.h:44-86
    method: str = "RTR"
    gradnorm_tol: float = 1e-2
    RGD_stepsize: float = 1e-3
    RGD_use_preconditioner: bool = True
    RTR_iterations: int = 3
    RTR_tCG_iterations: int = 50
    RTR_initial_radius: float = 100.0
    # ROPTLIB SolversTR defaults (not exposed by the reference)
    theta: float = 1.0
    kappa: float = 0.1
    accept_rho: float = 0.1
    shrink: float = 0.25
    magnify: float = 2.0


@dataclass
class ROPTResult:              # include/DPGO/DPGO_types.h:91-107
    success: bool = False
    fInit: float = 0.0
    gradNormInit: float = 0.0
    fOpt: float = 0.0
    gradNormOpt: float = 0.0
    tCGStatus: int = TCG_MAXITER
    # extras (oracle only)
    outer: int = 0
    inner_total: int = 0
    accepted: int = 0
    trace: list = field(default_factory=list)


def _metric(A, B):
    return float(np.sum(A * B))


def tcg(prob: QuadraticProblem, X, EG, grad, Delta, prm: ROptParameters):
    """SolversTR::tCG_TR (Steihaug-Toint, preconditioned, eta0 = 0)."""
    r = grad.copy()
    e_Pe = 0.0
    r_r = _metric(r, r)
    norm_r0 = math.sqrt(r_r)
    z = prob.precondition(X, r)
    z_r = _metric(z, r)
    d_Pd = z_r
    delta = -z
    e_Pd = 0.0
    eta = np.zeros_like(grad)
    status = TCG_MAXITER
    j = 0
    for j in range(prm.RTR_tCG_iterations):
        Hd = prob.rhess(X, EG, delta)
        d_Hd = _metric(delta, Hd)
        alpha = z_r / d_Hd if d_Hd != 0 else math.inf
        e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd
        if d_Hd <= 0 or e_Pe_new >= Delta * Delta:
            tau = (-e_Pd + math.sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd
            eta = eta + tau * delta
            status = TCG_NEGCURV if d_Hd <= 0 else TCG_EXCREGION
            break
        e_Pe = e_Pe_new
        eta = eta + alpha * delta
        r = r + alpha * Hd
        r_r = _metric(r, r)
        norm_r = math.sqrt(r_r)
        if norm_r <= norm_r0 * min(norm_r0 ** prm.theta, prm.kappa):
            status = TCG_LCON if prm.kappa < norm_r0 ** prm.theta else TCG_SCON
            break
        z = prob.precondition(X, r)
        zold_rold = z_r
        z_r = _metric(z, r)
        beta = z_r / zold_rold
        delta = -z + beta * delta
        e_Pd = beta * (e_Pd + alpha * d_Pd)
        d_Pd = z_r + beta * beta * d_Pd
    else:
        j = prm.RTR_tCG_iterations - 1
    inner = j + 1 if prm.RTR_tCG_iterations > 0 else 0
    return eta, status, inner


def rtr_run(prob: QuadraticProblem, X0, prm: ROptParameters, initial_delta, max_delta,
            max_iter, res: ROPTResult):
    """SolversTR::Run with Stop_Criterion = GRAD_F (src/QuadraticOptimizer.cpp:63-78)."""
    d = prob.d
    x1 = X0.copy()
    f1 = prob.f(x1)
    EG = prob.egrad(x1)
    gf1 = tangent_project(x1, EG, d)
    ngf = math.sqrt(_metric(gf1, gf1))
    Delta = initial_delta
    it = 0
    accepted_last = False
    stop = False
    status = TCG_MAXITER
    while (not stop) and it < max_iter:
        eta, status, inner = tcg(prob, x1, EG, gf1, Delta, prm)
        x2 = retract_qf(x1, eta, d)
        f2 = prob.f(x2)
        zeta = prob.rhess(x1, EG, eta)
        rho = (f1 - f2) / (-_metric(eta, gf1 + 0.5 * zeta))
        if rho > 0.75:
            if status in (TCG_EXCREGION, TCG_NEGCURV):
                Delta = min(prm.magnify * Delta, max_delta)
        elif rho < 0.25:
            Delta = prm.shrink * Delta
        sqeps = math.sqrt(np.finfo(float).eps)
        if rho > prm.accept_rho or (abs(f1 - f2) / (abs(f1) + 1) < sqeps and f2 < f1):
            EG2 = prob.egrad(x2)
            gf2 = tangent_project(x2, EG2, d)
            ngf = math.sqrt(_metric(gf2, gf2))
            x1, EG, gf1, f1 = x2, EG2, gf2, f2
            accepted_last = True
            res.accepted += 1
        else:
            accepted_last = False
        stop = ngf < prm.gradnorm_tol
        it += 1
        res.inner_total += inner
        res.trace.append(dict(iter=it, f=f1, gradnorm=ngf, rho=rho, Delta=Delta,
                              inner=inner, status=status, accepted=accepted_last))
    res.outer += it
    res.tCGStatus = status
    return x1, accepted_last


def trust_region(prob, Yinit, prm: ROptParameters, res: ROPTResult):
    """QuadraticOptimizer::trustRegion, src/QuadraticOptimizer.cpp:50-108."""
    gn0 = prob.rgrad_norm(Yinit)
    if gn0 < prm.gradnorm_tol:
        return Yinit.copy()
    if prm.RTR_iterations == 1:
        radius = prm.RTR_initial_radius
        total = 0
        while True:
            Y, acc = rtr_run(prob, Yinit, prm, radius, radius, 1, res)
            if acc:
                return Y
            if total > 10:
                return Yinit.copy()
            radius /= 4
            total += 1
    Y, _ = rtr_run(prob, Yinit, prm, prm.RTR_initial_radius, 5 * prm.RTR_initial_radius,
                   prm.RTR_iterations, res)
    return Y


def gradient_descent(prob, Yinit, prm: ROptParameters):
    """QuadraticOptimizer::gradientDescent, src/QuadraticOptimizer.cpp:110-137."""
    g = prob.rgrad(Yinit)
    if prm.RGD_use_preconditioner:
        g = prob.precondition(Yinit, g)
    return retract_qf(Yinit, -prm.RGD_stepsize * g, prob.d)


def optimize(prob, Y, prm: ROptParameters | None = None):
    """QuadraticOptimizer::optimize, src/QuadraticOptimizer.cpp:26-48."""
    prm = prm or ROptParameters()
    res = ROPTResult()
    res.fInit = prob.f(Y)
    res.gradNormInit = prob.rgrad_norm(Y)
    if prm.method == "RTR":
        Yopt = trust_region(prob, Y, prm, res)
    else:
        Yopt = gradient_descent(prob, Y, prm)
    res.fOpt = prob.f(Yopt)
    res.gradNormOpt = prob.rgrad_norm(Yopt)
    res.success = True
    return Yopt, res


# --------------------------------------------------------------------------------------
# initialization
# --------------------------------------------------------------------------------------
def measurement_error(R, t, kappa, tau, Y1, p1, Y2, p2):
    """computeMeasurementError, src/DPGO_utils.cpp:501-507 (squared, weighted by kappa / tau)."""
    rot = float(np.sum((Y1 @ R - Y2) ** 2))
    tr = float(np.sum((p2 - p1 - Y1 @ t) ** 2))
    return kappa * rot + tau * tr


class RobustCost:
    """RobustCost, src/DPGO_robust.cpp:45-134 (defaults: include/DPGO/DPGO_robust.h)."""

    def __init__(self, cost_type="GNC_TLS", gnc_max_iters=20, gnc_barc=5.0, gnc_mu_step=1.4, gnc_init_mu=1e-4,
                 huber_threshold=3.0, tls_threshold=10.0):
        self.type = cost_type
        self.gnc_max_iters, self.barc, self.mu_step, self.init_mu = gnc_max_iters, gnc_barc, gnc_mu_step, gnc_init_mu
        self.huber, self.tls = huber_threshold, tls_threshold
        self.reset()

    def reset(self):                      # :98-112
        self.mu = self.init_mu
        self.gnc_iteration = 0

    def weight(self, r):                  # :54-96
        if self.type == "L2":
            return 1.0
        if self.type == "L1":
            return 1.0 / r
        if self.type == "Huber":
            return 1.0 if r < self.huber else self.huber / r
        if self.type == "TLS":
            return 1.0 if r < self.tls else 0.0
        if self.type == "GM":
            a = 1 + r * r
            return 1.0 / (a * a)
        if self.type == "GNC_TLS":        # eq. (14) of the GNC paper
            r2, c2, mu = r * r, self.barc * self.barc, self.mu
            if r2 >= (mu + 1) / mu * c2:
                return 0.0
            if r2 <= mu / (mu + 1) * c2:
                return 1.0
            return float(np.sqrt(c2 * mu * (mu + 1) / r2) - mu)
        raise ValueError(self.type)

    def update(self):                     # :114-132
        if self.type != "GNC_TLS":
            return
        self.gnc_iteration += 1
        if self.gnc_iteration > self.gnc_max_iters:
            return
        self.mu *= self.mu_step


def single_rotation_averaging(Rs, kappa):
    """singleRotationAveraging, src/DPGO_solver.cpp:42-57."""
    return project_rotation(np.einsum("i,ijk->jk", kappa, Rs))


def _gnc_averaging(n, threshold, max_iters, solve, residual_sq):
    """The GNC-TLS loop of robustSingleRotationAveraging / robustSinglePoseAveraging
    (src/DPGO_solver.cpp:72-134, 136-218)."""
    w_tol = 1e-8
    w = np.ones(n)
    est = solve(w)
    rsq = np.array([residual_sq(est, i) for i in range(n)])
    mu_init = min(threshold ** 2 / (2 * rsq.max() - threshold ** 2), 1e-5)
    if mu_init > 0:
        cost = RobustCost("GNC_TLS", gnc_max_iters=max_iters, gnc_barc=threshold, gnc_init_mu=mu_init)
        for _ in range(max_iters):
            est = solve(w)
            w = np.array([cost.weight(math.sqrt(residual_sq(est, i))) for i in range(n)])
            if np.all((w < w_tol) | (w > 1 - w_tol)):
                break
            cost.update()
    return est, [i for i in range(n) if w[i] > 1 - w_tol]


def robust_single_rotation_averaging(Rs, kappa=None, threshold=0.1):
    Rs = np.asarray(Rs)
    kappa = np.ones(len(Rs)) if kappa is None else np.asarray(kappa, dtype=float)
    return _gnc_averaging(len(Rs), threshold, 1000, lambda w: single_rotation_averaging(Rs, kappa * w),
                          lambda R, i: kappa[i] * np.sum((R - Rs[i]) ** 2))


def robust_single_pose_averaging(Rs, ts, kappa=None, tau=None, threshold=0.1):
    Rs, ts = np.asarray(Rs), np.asarray(ts)
    n = len(Rs)
    kappa = 10000 * np.ones(n) if kappa is None else np.asarray(kappa, dtype=float)
    tau = 100 * np.ones(n) if tau is None else np.asarray(tau, dtype=float)

    def solve(w):                        # singlePoseAveraging :59-70
        return single_rotation_averaging(Rs, kappa * w), (tau * w) @ ts / np.sum(tau * w)

    def res(est, i):
        return kappa[i] * np.sum((est[0] - Rs[i]) ** 2) + tau[i] * np.sum((est[1] - ts[i]) ** 2)
    (R, t), inl = _gnc_averaging(n, threshold, 10000, solve, res)
    return R, t, inl


def odometry_initialization(odom: Measurements, n: int):
    """odometryInitialization, src/DPGO_solver.cpp:271-303 (identity start)."""
    d = odom.d
    T = np.zeros((d, (d + 1) * n))
    T[:, :d] = np.eye(d)
    order = np.argsort(odom.p1)
    for k in order:
        src, dst = int(odom.p1[k]), int(odom.p2[k])
        assert dst == src + 1
        Rs = T[:, src * (d + 1):src * (d + 1) + d]
        ts = T[:, src * (d + 1) + d]
        T[:, dst * (d + 1):dst * (d + 1) + d] = Rs @ odom.R[k]
        T[:, dst * (d + 1) + d] = ts + Rs @ odom.t[k]
    return T


def chordal_initialization(meas: Measurements, n: int):
    """chordalInitialization, src/DPGO_solver.cpp:220-269 + recoverTranslations
    src/DPGO_utils.cpp:435-462.  The reference solves the two sparse least-squares problems
    with SPQR; the minimizers are unique (pose 0 anchored) so normal equations give the same
    answer up to rounding."""
    d = meas.d
    m = len(meas)
    # rotations: min sum kappa || R_j - R_i R_ij ||_F^2, R_0 = I.  Row-wise independent:
    # each row rho of R_i is a 1 x d vector x_i with x_j - x_i R_ij = 0.
    rows, cols, vals = [], [], []
    for e in range(m):
        i, j = int(meas.p1[e]), int(meas.p2[e])
        sk = math.sqrt(meas.kappa[e])
        for c in range(d):
            for a in range(d):
                rows.append(e * d + c); cols.append(i * d + a); vals.append(-sk * meas.R[e][a, c])
            rows.append(e * d + c); cols.append(j * d + c); vals.append(sk)
    B = sp.coo_matrix((vals, (rows, cols)), shape=(m * d, n * d)).tocsc()
    Bred = B[:, d:]
    rhs = -(B[:, :d] @ np.eye(d))                 # anchored R_0 = I, one rhs column per row rho
    N = (Bred.T @ Bred).tocsc()
    lu = spla.splu(N)
    sol = lu.solve(np.asarray(Bred.T @ rhs))       # (d(n-1), d): column rho = row rho of all R_i
    Rch = np.zeros((d, d * n))
    Rch[:, :d] = np.eye(d)
    for i in range(1, n):
        blk = sol[(i - 1) * d:i * d, :].T          # rows rho, cols a
        Rch[:, i * d:(i + 1) * d] = project_rotation(blk)
    # translations: min sum tau || t_j - t_i - R_i t_ij ||^2, t_0 = 0
    rows, cols, vals = [], [], []
    c = np.zeros((m, d))
    for e in range(m):
        i, j = int(meas.p1[e]), int(meas.p2[e])
        st = math.sqrt(meas.tau[e])
        rows += [e, e]; cols += [i, j]; vals += [-st, st]
        c[e] = st * (Rch[:, i * d:(i + 1) * d] @ meas.t[e])
    B1 = sp.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsc()
    B1r = B1[:, 1:]
    lu = spla.splu((B1r.T @ B1r).tocsc())
    tred = lu.solve(np.asarray(B1r.T @ c))        # (n-1, d)
    T = np.zeros((d, (d + 1) * n))
    for i in range(n):
        T[:, i * (d + 1):i * (d + 1) + d] = Rch[:, i * d:(i + 1) * d]
        if i > 0:
            T[:, i * (d + 1) + d] = tred[i - 1]
    return T


def lifting_matrix(d, r):
    """Stand-in for fixedStiefelVariable (src/DPGO_utils.cpp:488-493), which depends on
    srand(1) + ROPTLIB's RNG and cannot be reproduced.  Any fixed element of St(d, r) gives the
    same objective sequence (f is invariant under a common left-orthogonal transform), so the
    oracle and the product both use the deterministic QF of a fixed full-rank pattern."""
    A = np.zeros((r, d))
    for i in range(r):
        for j in range(d):
            A[i, j] = math.cos(1.0 + 0.7 * i + 1.3 * j) + (1.0 if i == j else 0.0)
    Qm, Rm = np.linalg.qr(A)
    sg = np.sign(np.diag(Rm)); sg[sg == 0] = 1
    return Qm * sg


def round_solution(X, d):
    """PGOAgent::getTrajectoryInLocalFrame, src/PGOAgent.cpp:718-736."""
    Y0 = X[:, :d]
    T = Y0.T @ X
    n = X.shape[1] // (d + 1)
    t0 = T[:, d].copy()
    for i in range(n):
        T[:, i * (d + 1):i * (d + 1) + d] = project_rotation(T[:, i * (d + 1):i * (d + 1) + d])
        T[:, i * (d + 1) + d] -= t0
    return T
