"""CPU test of the persistent fused RTR solver kernel itself: k_rtr_fused<R, D, MODE>
(dpgo_b200/csrc/fused_kernel.cuh, the source nvcc compiles for the product) is built with g++ against
tests/native/cuda_emu.h -- one CTA of 256 real threads, CTA barriers for __syncthreads and grid.sync, warp
shuffles, TMA bulk copies completing emulated mbarriers -- and must reproduce the oracle's
QuadraticOptimizer::optimize (ref: src/QuadraticOptimizer.cpp:26-108): same outer / tCG iteration counts,
same objective, same iterate.  MODE 0 = dense inverse (the default below 3000 scalars), MODE 2 = the
five-phase two-level preconditioner (the default above, i.e. the kernel behind the sphere2500 bench line),
MODE 3 / 4 = its three-phase forms, whose first run on a device is still pending.
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

import three_phase_emu as emu  # noqa: E402
from oracle import pgo  # noqa: E402
_dp, _ip, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p


class EmuProblem(C.Structure):
    _fields_ = ([("mode", C.c_int), ("R", C.c_int), ("d", C.c_int), ("n", C.c_int),
                 ("rowptr", _ip), ("colidx", _ip), ("blocks", _dp), ("G", _dp),
                 ("Pinv", _dp), ("ld", C.c_int), ("KT", C.c_int), ("nsplit", C.c_int)] +
                [(k, C.c_int) for k in ("V", "nS", "nsplit3", "sep_col0", "pcols", "ycols", "prefetch")] +
                [("M1", _dp), ("M3", _dp), ("M5", _dp), ("strips1", _vp), ("strips3", _vp), ("strips5", _vp)] +
                [(k, _ip) for k in ("cta1", "chunks1", "cta3", "chunks3", "cta5", "chunks5",
                                    "gidx", "icol", "tptr", "tcol", "pcol", "srow")] +
                [(k, _ip) for k in ("si_rowptr", "si_colidx", "bs_rowptr", "bs_colidx", "bcol")] +
                [("si_blocks", _dp), ("bs_blocks", _dp), ("nB", C.c_int)] +
                [(k, C.c_double) for k in ("gradnorm_tol", "init_radius", "theta", "kappa", "accept_rho", "shrink",
                                           "magnify")] +
                [("max_outer", C.c_int), ("max_inner", C.c_int), ("x_in", _dp), ("x_out", _dp), ("result", _dp)])


STRIP_DT = np.dtype([("cb", np.int32), ("kc0", np.int32), ("nchunks", np.int32), ("slot", np.int32),
                     ("data_off", np.int64)])


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
    else:
        from dpgo_b200 import _lib
        G = sp.coo_matrix((np.ones(len(meas.p1)), (meas.p1, meas.p2)), shape=(n, n))
        G = (G + G.T + sp.identity(n)).tocsr()
        G.sort_indices()
        plan = emu.fetch_plan(_lib.lib.dpgo_three_phase_plan, n, G.indptr, G.indices, dh, max_poses, V)
        bufs = emu.fill_stage_buffers(plan, *emu.dense_blocks(A, plan))
        for ph in ("1", "3", "5"):
            st = plan["strips" + ph]
            a = hold(np.zeros(max(len(st), 1), dtype=STRIP_DT))
            for k, name in enumerate(("cb", "kc0", "nchunks", "slot", "data_off")):
                a[name][:len(st)] = st[:, k]
            cta = hold(np.ascontiguousarray(plan["cta" + ph], dtype=np.int32))
            chunks = hold(np.array([st[cta[v]:cta[v + 1], 2].sum() for v in range(plan["V"])], dtype=np.int32))
            M = hold(np.ascontiguousarray(bufs[ph].reshape(-1)) if bufs[ph].size else np.zeros(1))
            setattr(e, "M" + ph, _dp_of(M))
            setattr(e, "strips" + ph, a.ctypes.data_as(_vp))
            setattr(e, "cta" + ph, _ip_of(cta))
            setattr(e, "chunks" + ph, _ip_of(chunks))
        for name in ("gidx", "icol", "tptr", "tcol", "pcol", "srow"):
            arr = np.ascontiguousarray(plan[name], dtype=np.int32)
            setattr(e, name, _ip_of(hold(arr if len(arr) else np.zeros(1, dtype=np.int32))))
        for name in ("V", "nS", "nsplit3", "sep_col0", "pcols", "ycols"):
            setattr(e, name, plan[name])
        e.prefetch = prefetch
        if mode == 2:
            _five_phase(e, plan, B, dh, hold)
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


def _five_phase(e, plan, B, dh, hold):
    """The five-phase form (MODE 2) on the same partition and column order: the interior strips are the M_k
    strips of the three-phase plan, the couplings stay sparse -- A_SI (rows = separator positions) and A_BS
    (rows = interior poses with a separator neighbour) as block-CSR with permuted scalar columns, which is
    what the device set-up (dd_build) uploads."""
    group, pcol, srow = plan["group"], plan["pcol"], plan["srow"]
    st, cta = plan["strips1"], plan["cta1"]
    keep_rows, new_cta = [], [0]
    for v in range(plan["V"]):
        keep_rows += [i for i in range(cta[v], cta[v + 1]) if st[i, 5] == 0]
        new_cta.append(len(keep_rows))
    a = hold(np.zeros(max(len(keep_rows), 1), dtype=STRIP_DT))
    for k, name in enumerate(("cb", "kc0", "nchunks", "slot", "data_off")):
        a[name][:len(keep_rows)] = st[keep_rows, k]
    ncta = hold(np.array(new_cta, dtype=np.int32))
    chunks = hold(np.array([a["nchunks"][ncta[v]:ncta[v + 1]].sum() for v in range(plan["V"])], dtype=np.int32))
    e.strips1, e.cta1, e.chunks1 = a.ctypes.data_as(_vp), _ip_of(ncta), _ip_of(chunks)

    def rows(poses, want_sep_cols):
        rp, ci, bl, used = [0], [], [], []
        for i in poses:
            n0 = len(ci)
            for p in range(B.indptr[i], B.indptr[i + 1]):
                c = B.indices[p]
                if (group[c] < 0) == want_sep_cols:
                    ci.append(pcol[c]); bl.append(B.data[p])
            if want_sep_cols and len(ci) == n0:
                continue                                     # interior pose without a separator neighbour: no row
            rp.append(len(ci)); used.append(i)
        blocks = np.ascontiguousarray(np.array(bl).reshape(-1)) if bl else np.zeros(1)
        return (hold(np.array(rp, dtype=np.int32)), hold(np.array(ci if ci else [0], dtype=np.int32)), hold(blocks), used)

    e.si_rowptr, e.si_colidx, e.si_blocks = (lambda r: (_ip_of(r[0]), _ip_of(r[1]), _dp_of(r[2])))(rows(srow, False))
    interior = [i for k in range(plan["K"]) for i in np.where(group == k)[0]]
    rp, ci, bl, used = rows(interior, True)
    e.bs_rowptr, e.bs_colidx, e.bs_blocks = _ip_of(rp), _ip_of(ci), _dp_of(bl)
    e.bcol = _ip_of(hold(np.array([pcol[i] for i in used] or [0], dtype=np.int32)))
    e.nB = len(used)


def _dp_of(a):
    return a.ctypes.data_as(_dp)


def _ip_of(a):
    return a.ctypes.data_as(_ip)


CASES = [
    ("tinyGrid3D", 5, 0, 0, 0, 1),
    ("smallGrid3D", 5, 0, 0, 0, 1),          # the product's default path for this size (dense inverse)
    ("smallGrid3D", 5, 2, 12, 5, 1),         # five-phase form of the two-level preconditioner (sphere2500's default)
    ("smallGrid3D", 5, 2, 60, 3, 0),         # the same with multi-wave strips, no prefetch
    ("smallGrid3D", 5, 3, 12, 5, 1),         # three-phase form, separate finish
    ("smallGrid3D", 5, 4, 12, 5, 1),         # finish fused into the last strip phase
    ("smallGrid3D", 5, 4, 12, 5, 0),         # the same without the pre-barrier prefetch
    ("smallGrid3D", 3, 4, 60, 3, 1),         # r = d, domains longer than one wave
    ("smallGrid3D", 5, 4, 200, 2, 1),        # a single domain, no separator
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
