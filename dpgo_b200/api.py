"""Thin numpy-facing wrapper over the C-ABI (include/dpgo_b200.h) used by the tests and by
bench.py.  The production host side is the C++ drop-in under dpgo_b200/host (PoseGraph /
QuadraticProblem / QuadraticOptimizer / PGOAgent shells); this module only marshals arrays.

Names mirror the reference: `f`, `RieGrad`, `RieGradNorm` (src/QuadraticProblem.cpp:29-83),
`optimize` / `getOptResult` (src/QuadraticOptimizer.cpp:26-48).
"""
import ctypes as C

import numpy as np

from ._lib import ChordalInfo, RoptParams, RoptResult, check, lib

SLOT_X, SLOT_Y, SLOT_V, SLOT_XPREV = 0, 1, 2, 3
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def default_params(**kw):
    p = RoptParams()
    lib.dpgo_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class DeviceProblem:
    """One agent's pose graph + quadratic problem + optimizer state on one GPU."""

    def __init__(self, n, d, r, device=0, stream=None):
        self.n, self.d, self.r = int(n), int(d), int(r)
        self.N = (self.d + 1) * self.n
        self._h = C.c_void_p()
        check(lib.dpgo_create(int(device), self.n, self.d, self.r,
                              C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.last_result = None
        self.num_private_edges = self.num_shared_edges = 0

    def close(self):
        if self._h:
            lib.dpgo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- graph -----------------------------------------------------------------------------
    @staticmethod
    def _edge_arrays(d, R, t, kappa, tau, weight, m):
        R = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(m, d, d))
        t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(m, d))
        kappa = np.ascontiguousarray(kappa, dtype=np.float64)
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        weight = np.ones(m) if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
        return R, t, kappa, tau, weight

    def set_private_edges(self, p1, p2, R, t, kappa, tau, weight=None):
        m = len(p1)
        p1 = np.ascontiguousarray(p1, dtype=np.int32)
        p2 = np.ascontiguousarray(p2, dtype=np.int32)
        R, t, kappa, tau, weight = self._edge_arrays(self.d, R, t, kappa, tau, weight, m)
        check(lib.dpgo_set_private_edges(self._h, m, _i(p1), _i(p2), _d(R), _d(t), _d(kappa), _d(tau),
                                         _d(weight)))
        self.num_private_edges = m

    def set_shared_edges(self, my_idx, nbr_slot, outgoing, R, t, kappa, tau, num_nbr_slots, weight=None):
        m = len(my_idx)
        my_idx = np.ascontiguousarray(my_idx, dtype=np.int32)
        nbr_slot = np.ascontiguousarray(nbr_slot, dtype=np.int32)
        outgoing = np.ascontiguousarray(outgoing, dtype=np.uint8)
        R, t, kappa, tau, weight = self._edge_arrays(self.d, R, t, kappa, tau, weight, m)
        check(lib.dpgo_set_shared_edges(self._h, m, int(num_nbr_slots), _i(my_idx), _i(nbr_slot),
                                        outgoing.ctypes.data_as(_bp), _d(R), _d(t), _d(kappa),
                                        _d(tau), _d(weight)))
        self.num_shared_edges = m

    def set_priors(self, idx, poses, prior_kappa=10000.0, prior_tau=100.0):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        tiles = np.ascontiguousarray(
            np.stack([np.asfortranarray(p, dtype=np.float64).ravel(order="F") for p in poses])
            if len(idx) else np.zeros((0, self.r * (self.d + 1))))
        check(lib.dpgo_set_priors(self._h, len(idx), _i(idx), _d(tiles), prior_kappa, prior_tau))

    def set_precon_mode(self, mode):
        """-1 = by size (default), 0 = full dense inverse, 1 = symmetric half storage,
        2 = two-level (nested-dissection domains + separator Schur complement)."""
        check(lib.dpgo_set_precon_mode(self._h, int(mode)))

    def precon_mode(self):
        v = C.c_int()
        check(lib.dpgo_get_precon_mode(self._h, C.byref(v)))
        return v.value

    def set_precon_tuning(self, split_interior=0, split_schur=0, prefetch=-1):
        check(lib.dpgo_set_precon_tuning(self._h, int(split_interior), int(split_schur), int(prefetch)))

    def round_trajectory(self, slot=SLOT_X, anchor=None):
        """d x (d+1)n rounded poses of the slot in the frame of `anchor` (r x (d+1) lifted pose; None: pose 0)."""
        out = np.zeros((self.d, (self.d + 1) * self.n), order="F")
        a = None if anchor is None else np.asfortranarray(anchor, dtype=np.float64)
        check(lib.dpgo_round_trajectory(self._h, int(slot), _d(a) if a is not None else None, _d(out)))
        return out

    def neighbor_buffer(self, aux):
        """Device address of the handle's neighbour pose buffer (aux = 0: X, 1: auxiliary Y)."""
        p = C.c_void_p()
        check(lib.dpgo_neighbor_buffer(self._h, int(aux), C.byref(p)))
        return p.value

    def use_neighbor_poses(self, aux):
        check(lib.dpgo_use_neighbor_poses(self._h, int(aux)))

    def set_qx_variant(self, variant=0, prefetch_distance=0):
        check(lib.dpgo_set_qx_variant(self._h, int(variant), int(prefetch_distance)))

    def set_two_level_domain_size(self, max_domain_poses=0):
        check(lib.dpgo_set_two_level_domain_size(self._h, int(max_domain_poses)))

    def finalize(self, build_precon=True):
        check(lib.dpgo_finalize(self._h, 1 if build_precon else 0))

    def update_weights(self, w_private=None, w_shared=None, build_precon=True):
        wp = None if w_private is None else np.ascontiguousarray(w_private, dtype=np.float64)
        ws = None if w_shared is None else np.ascontiguousarray(w_shared, dtype=np.float64)
        check(lib.dpgo_update_weights(self._h, _d(wp) if wp is not None else None,
                                      _d(ws) if ws is not None else None, 1 if build_precon else 0))

    def get_Q_bsr(self):
        nnzb = C.c_int()
        check(lib.dpgo_get_Q_bsr(self._h, C.byref(nnzb), None, None, None))
        dh = self.d + 1
        rowptr = np.zeros(self.n + 1, dtype=np.int32)
        colidx = np.zeros(nnzb.value, dtype=np.int32)
        blocks = np.zeros((nnzb.value, dh, dh))
        check(lib.dpgo_get_Q_bsr(self._h, C.byref(nnzb), _i(rowptr), _i(colidx), _d(blocks)))
        return rowptr, colidx, blocks

    # ---- arrays ----------------------------------------------------------------------------
    def _in(self, X):
        X = np.asfortranarray(X, dtype=np.float64)
        if X.shape != (self.r, self.N):
            raise ValueError(f"expected {(self.r, self.N)}, got {X.shape}")
        return X

    def _out(self):
        return np.empty((self.r, self.N), dtype=np.float64, order="F")

    def set_G(self, G):
        G = self._in(G)
        check(lib.dpgo_set_G(self._h, _d(G)))

    def get_G(self):
        out = self._out()
        check(lib.dpgo_get_G(self._h, _d(out)))
        return out

    def set_neighbor_poses(self, tiles):
        """tiles: (num_nbr_slots, r, d+1) array of neighbour poses (slot order)."""
        t = np.ascontiguousarray(np.asarray(tiles, dtype=np.float64).transpose(0, 2, 1))
        check(lib.dpgo_set_neighbor_poses(self._h, _d(t)))

    def set_neighbor_poses_dev(self, dev_ptr):
        check(lib.dpgo_set_neighbor_poses_dev(self._h, C.c_void_p(dev_ptr)))

    # ---- QuadraticProblem ------------------------------------------------------------------
    def qx(self, X):
        X = self._in(X); out = self._out()
        check(lib.dpgo_qx(self._h, _d(X), _d(out)))
        return out

    def f(self, X):
        X = self._in(X); v = C.c_double()
        check(lib.dpgo_f(self._h, _d(X), C.byref(v)))
        return v.value

    def egrad(self, X):
        X = self._in(X); out = self._out()
        check(lib.dpgo_egrad(self._h, _d(X), _d(out)))
        return out

    def RieGrad(self, X):
        X = self._in(X); out = self._out(); v = C.c_double()
        check(lib.dpgo_rgrad(self._h, _d(X), _d(out), C.byref(v)))
        return out

    def RieGradNorm(self, X):
        X = self._in(X); v = C.c_double()
        check(lib.dpgo_rgrad(self._h, _d(X), None, C.byref(v)))
        return v.value

    def hessvec(self, X, V):
        X = self._in(X); V = self._in(V); out = self._out()
        check(lib.dpgo_hessvec(self._h, _d(X), _d(V), _d(out)))
        return out

    def precon(self, X, V):
        X = self._in(X); V = self._in(V); out = self._out()
        check(lib.dpgo_precon(self._h, _d(X), _d(V), _d(out)))
        return out

    def tangent_project(self, X, V):
        X = self._in(X); V = self._in(V); out = self._out()
        check(lib.dpgo_tangent_project(self._h, _d(X), _d(V), _d(out)))
        return out

    def retract(self, X, V):
        X = self._in(X); V = self._in(V); out = self._out()
        check(lib.dpgo_retract(self._h, _d(X), _d(V), _d(out)))
        return out

    def project_manifold(self, M):
        M = self._in(M); out = self._out()
        check(lib.dpgo_project_manifold(self._h, _d(M), _d(out)))
        return out

    # ---- QuadraticOptimizer ----------------------------------------------------------------
    def optimize(self, X0=None, params=None, fetch=True):
        res = RoptResult()
        X0 = None if X0 is None else self._in(X0)
        out = self._out() if fetch else None
        check(lib.dpgo_optimize(self._h, C.byref(params) if params is not None else None,
                                _d(X0) if X0 is not None else None,
                                _d(out) if out is not None else None, C.byref(res)))
        self.last_result = res.as_dict()
        return out, self.last_result

    def getOptResult(self):
        return self.last_result

    # ---- agent state -----------------------------------------------------------------------
    def slot_set(self, slot, X):
        X = self._in(X)
        check(lib.dpgo_slot_set(self._h, slot, _d(X)))

    def slot_get(self, slot):
        out = self._out()
        check(lib.dpgo_slot_get(self._h, slot, _d(out)))
        return out

    def slot_copy(self, dst, src):
        check(lib.dpgo_slot_copy(self._h, dst, src))

    def nesterov_update_Y(self, alpha):
        check(lib.dpgo_nesterov_update_Y(self._h, float(alpha)))

    def nesterov_update_V(self, gamma):
        check(lib.dpgo_nesterov_update_V(self._h, float(gamma)))

    def optimize_slot(self, src_slot, params=None):
        res = RoptResult()
        check(lib.dpgo_optimize_slot(self._h, C.byref(params) if params is not None else None,
                                     int(src_slot), C.byref(res)))
        self.last_result = res.as_dict()
        return self.last_result

    def optimize_slot_async(self, src_slot, params=None):
        """Queue the fused solve on the handle's stream and return without waiting."""
        check(lib.dpgo_optimize_slot_async(self._h, C.byref(params) if params is not None else None,
                                           int(src_slot)))

    def optimize_result(self):
        """Wait for the stream; result block of the most recent optimize_slot_async."""
        res = RoptResult()
        check(lib.dpgo_optimize_result(self._h, C.byref(res)))
        self.last_result = res.as_dict()
        return self.last_result

    def set_public_indices(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        check(lib.dpgo_set_public_indices(self._h, len(idx), _i(idx)))

    def pack_public_dev(self, slot, dev_ptr):
        check(lib.dpgo_pack_public_dev(self._h, int(slot), C.c_void_p(dev_ptr)))

    def gather_tiles_dev(self, slot, num, idx_dev_ptr, out_dev_ptr):
        check(lib.dpgo_gather_tiles_dev(self._h, int(slot), int(num), C.c_void_p(idx_dev_ptr),
                                        C.c_void_p(out_dev_ptr)))

    def measurement_errors(self, slot, nbr_dev_ptr=None):
        """Squared residual of every private / shared edge at the poses in `slot`
        (computeMeasurementError, src/DPGO_utils.cpp:501-507), computed on the device."""
        ep = np.zeros(max(self.num_private_edges, 1))
        es = np.zeros(max(self.num_shared_edges, 1))
        check(lib.dpgo_measurement_errors(self._h, int(slot), C.c_void_p(nbr_dev_ptr or 0), _d(ep), _d(es)))
        return ep[:self.num_private_edges], es[:self.num_shared_edges]

    def max_translation_distance(self, a, b):
        v = C.c_double()
        check(lib.dpgo_max_translation_distance(self._h, a, b, C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_int64()
        check(lib.dpgo_launch_count(self._h, C.byref(v)))
        return v.value

    def sync(self):
        check(lib.dpgo_sync(self._h))

    # ---- measurement -----------------------------------------------------------------------
    def time_qx(self, reps=20, flush_l2=False):
        v = C.c_double()
        check(lib.dpgo_time_qx(self._h, reps, 1 if flush_l2 else 0, C.byref(v)))
        return v.value

    def time_pose_op(self, op, reps=20, flush_l2=False):
        """op 0 = QF retraction, 1 = polar projection (Nesterov form), 2 = rounding; microseconds per launch"""
        v = C.c_double()
        check(lib.dpgo_time_pose_op(self._h, int(op), reps, 1 if flush_l2 else 0, C.byref(v)))
        return v.value

    def time_precon(self, reps=10, flush_l2=False):
        v = C.c_double()
        check(lib.dpgo_time_precon(self._h, reps, 1 if flush_l2 else 0, C.byref(v)))
        return v.value

    def bytes_qx(self):
        v = C.c_double()
        check(lib.dpgo_bytes_qx(self._h, C.byref(v)))
        return v.value

    def bytes_precon(self):
        v = C.c_double()
        check(lib.dpgo_bytes_precon(self._h, C.byref(v)))
        return v.value


def problem_from_measurements(p1, p2, R, t, kappa, tau, n, d, r, device=0, stream=None,
                              build_precon=True, weight=None, precon_mode=None, precon_tuning=None,
                              domain_size=None):
    """Single-robot problem (all edges private), as examples/MultiRobotExample.cpp:61-63 builds
    `problemCentral`."""
    prob = DeviceProblem(n, d, r, device, stream)
    prob.set_private_edges(p1, p2, R, t, kappa, tau, weight)
    if precon_mode is not None:
        prob.set_precon_mode(precon_mode)
    if precon_tuning is not None:
        prob.set_precon_tuning(*precon_tuning)
    if domain_size is not None:
        prob.set_two_level_domain_size(domain_size)
    prob.finalize(build_precon)
    return prob


def chordal_initialization(p1, p2, R, t, kappa, tau, n, d, device=0):
    """chordalInitialization (src/DPGO_solver.cpp:220-269) on the device; returns (T, info) with T the
    d x (d+1)n pose array and info the iteration counts / relative residuals of the two linear solves."""
    m = len(p1)
    p1 = np.ascontiguousarray(p1, dtype=np.int32)
    p2 = np.ascontiguousarray(p2, dtype=np.int32)
    R, t, kappa, tau, _ = DeviceProblem._edge_arrays(d, R, t, kappa, tau, None, m)
    T = np.zeros((d, (d + 1) * n), order="F")
    info = ChordalInfo()
    check(lib.dpgo_chordal_initialization(device, n, d, m, _i(p1), _i(p2), _d(R), _d(t), _d(kappa), _d(tau), _d(T),
                                          C.byref(info)))
    return T, {k: getattr(info, k) for k, _ in ChordalInfo._fields_}
