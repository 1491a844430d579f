"""CPU test of the persistent fused RTR solver kernel itself: k_rtr_fused<R, D, 0>
(dpgo_b200/csrc/fused_kernel.cuh, the source nvcc compiles for the product) is built with g++ against
tests/native/cuda_emu.h -- one CTA of 256 real threads, CTA barriers for __syncthreads and grid.sync, warp
shuffles, TMA bulk copies completing emulated mbarriers -- and must reproduce the oracle's
QuadraticOptimizer::optimize (ref: src/QuadraticOptimizer.cpp:26-108): same outer / tCG iteration counts,
same objective, same iterate.  MODE 0 = dense inverse (the default below 3000 scalars); the two-level form
(MODE 2) is covered on the device by tests/test_gpu_a_parity.py.
What the emulation cannot show: timing, the asynchrony of real TMA copies, memory-model effects across SMs."""
import ctypes as C
import os
import subprocess

import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):       # also when run as a script (one case per process)
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import pgo  # noqa: E402
_dp, _ip, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p


class EmuProblem(C.Structure):
    _fields_ = ([("mode", C.c_int), ("R", C.c_int), ("d", C.c_int), ("n", C.c_int),
                 ("rowptr", _ip), ("colidx", _ip), ("blocks", _dp), ("G", _dp),
                 ("Pinv", _dp), ("ld", C.c_int), ("KT", C.c_int), ("nsplit", C.c_int)] +
                [(k, C.c_double) for k in ("gradnorm_tol", "init_radius", "theta", "kappa", "accept_rho", "shrink",
                                           "magnify")] +
                [("max_outer", C.c_int), ("max_inner", C.c_int), ("x_in", _dp), ("x_out", _dp), ("result", _dp)])


@pytest.fixture(scope="module")
def solver_emu_so(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("native") / "libfused_solver_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                           "-fvisibility=hidden", "-fno-gnu-unique", "-Wl,-Bsymbolic",
                           os.path.join(ROOT, "tests", "native", "fused_solver_emu.cpp"), "-o", so])
    return so


def _load(so):
    lib = C.CDLL(so)
    lib.fused_solve_emu.restype = C.c_int
    lib.fused_solve_emu.argtypes = [C.POINTER(EmuProblem)]
    return lib


def _dense_tiling(N, num_sms=148):
    """ld, KT, nsplit as dpgo_create chooses them (dpgo_b200/csrc/device_lib.cu)."""
    ld = -(-N // 128) * 128
    ncb, wave = ld // 64, num_sms * 2
    best, KT_best, ns_best = -1.0, 0, 0
    for ns in range(6, 25):
        KT = -(-(-(-ld // ns)) // 32) * 32
        KT = max(KT, 256)
        nsp = -(-ld // KT)
        tiles = ncb * nsp
        rounds = -(-tiles // wave)
        eff = (tiles / (rounds * wave)) * (ld / (nsp * KT))
        score = 0.5 * eff if tiles < wave else eff
        if score > best + 1e-9:
            best, KT_best, ns_best = score, KT, nsp
    return ld, KT_best, ns_best


def _solve(lib, meas, n, R, mode, X0, max_poses=12, V=5, prefetch=1):
    d, dh = meas.d, meas.d + 1
    N = dh * n
    keep = []
    hold = lambda a: (keep.append(a), a)[1]
    Q = pgo.connection_laplacian(meas, n)
    B = sp.bsr_matrix(Q, blocksize=(dh, dh))
    B.sort_indices()
    e = EmuProblem()
    e.mode, e.R, e.d, e.n = mode, R, d, n
    e.rowptr = _ip_of(hold(np.ascontiguousarray(B.indptr, dtype=np.int32)))
    e.colidx = _ip_of(hold(np.ascontiguousarray(B.indices, dtype=np.int32)))
    e.blocks = _dp_of(hold(np.ascontiguousarray(B.data, dtype=np.float64)))
    e.G = _dp_of(hold(np.zeros(R * N)))
    A = (Q + 0.1 * sp.identity(N)).tocsc()
    if mode == 0:
        ld, KT, nsplit = _dense_tiling(N)
        ldk = KT * nsplit
        P = np.zeros((ld, ldk))
        P[:N, :N] = np.linalg.inv(A.toarray())
        ncb = ld // 64
        # T[((kc * ncb + cb) * 32 + kk) * 64 + jj] = P[cb * 64 + jj, kc * 32 + kk]
        T = P.reshape(ncb, 64, ldk // 32, 32).transpose(2, 0, 3, 1)
        e.Pinv = _dp_of(hold(np.ascontiguousarray(T).reshape(-1)))
        e.ld, e.KT, e.nsplit = ld, KT, nsplit
    prm = pgo.ROptParameters()
    e.gradnorm_tol, e.init_radius = prm.gradnorm_tol, prm.RTR_initial_radius
    e.theta, e.kappa, e.accept_rho, e.shrink, e.magnify = 1.0, 0.1, 0.1, 0.25, 2.0     # ROPTLIB defaults (rtr_logic.h)
    e.max_outer, e.max_inner = prm.RTR_iterations, prm.RTR_tCG_iterations
    xin = hold(np.ascontiguousarray(X0.T).reshape(-1))
    xout = hold(np.zeros(R * N))
    res = hold(np.zeros(16))
    e.x_in, e.x_out, e.result = _dp_of(xin), _dp_of(xout), _dp_of(res)
    assert lib.fused_solve_emu(C.byref(e)) == 0
    names = ["f_init", "gn_init", "f_opt", "gn_opt", "outer", "inner", "accepted", "rejected", "tcg_status",
             "returned_initial", "n_qx", "n_precon", "n_sweeps", "n_barriers"]
    return xout.reshape(N, R).T.copy(), dict(zip(names, res[:14]))


def _dp_of(a):
    return a.ctypes.data_as(_dp)


def _ip_of(a):
    return a.ctypes.data_as(_ip)


CASES = [
    ("tinyGrid3D", 5, 0, 0, 0, 1),
    ("smallGrid3D", 5, 0, 0, 0, 1),          # the product's default path for this size (dense inverse)
    ("smallGrid3D", 3, 0, 0, 0, 1),          # r = d
]


def _run_case(so, name, R, mode, max_poses, V, prefetch):
    from conftest import load_dataset
    meas, n, z = load_dataset(name)
    d = meas.d
    X0 = pgo.lifting_matrix(d, R) @ z["T_chordal"]
    Xg, res = _solve(_load(so), meas, n, R, mode, X0, max_poses, V, prefetch)
    op = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((R, (d + 1) * n)), d)
    Xo, ro = pgo.optimize(op, X0)
    assert (int(res["outer"]), int(res["inner"])) == (ro.outer, ro.inner_total), (res, ro.outer, ro.inner_total)
    assert abs(res["f_init"] - ro.fInit) <= 1e-10 * abs(ro.fInit)
    assert abs(res["f_opt"] - ro.fOpt) <= 1e-9 * abs(ro.fOpt)
    assert np.linalg.norm(Xg - Xo) <= 1e-6 * np.linalg.norm(Xo)
    assert res["n_precon"] >= res["inner"] and res["n_barriers"] > 0


@pytest.mark.parametrize("name,R,mode,max_poses,V,prefetch", CASES)
def test_fused_solver_kernel_matches_the_oracle(solver_emu_so, name, R, mode, max_poses, V, prefetch):
    """Each case in a fresh interpreter: the emulation runs 256 OS threads that meet at barriers thousands of
    times, and must not share a process with whatever thread pools earlier tests have started."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.abspath(__file__),
                          json.dumps([solver_emu_so, name, R, mode, max_poses, V, prefetch])],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


if __name__ == "__main__":
    import json
    import sys
    _run_case(*json.loads(sys.argv[1]))
