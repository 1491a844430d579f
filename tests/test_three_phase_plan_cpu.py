"""CPU tests of the three-phase form of the two-level preconditioner (precon_mode 3): the index plan
the library builds on the host (dpgo_b200/csrc/dd_plan.h through dpgo_three_phase_plan) is replayed
in numpy (tests/three_phase_emu.py: stage buffers filled as the device set-up lays them out, the
three strip phases with the kernels' staging rules, the finish) and must reproduce the exact
(Q + 0.1 I)^-1 -- the reference's preconditioner (src/QuadraticProblem.cpp:56-69).  What this does
not cover is the CUDA code itself; that is tests/test_gpu_zzz_three_phase.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):       # also when run as a script (one emulation case per process)
    if _p not in sys.path:
        sys.path.insert(0, _p)

import three_phase_emu as emu  # noqa: E402
from oracle import pgo  # noqa: E402


def _fn():
    from dpgo_b200 import _lib
    return _lib.lib.dpgo_three_phase_plan


def _pose_graph(p1, p2, n):
    A = sp.coo_matrix((np.ones(len(p1)), (p1, p2)), shape=(n, n))
    A = (A + A.T + sp.identity(n)).tocsr()
    A.sort_indices()
    return A


def _replay(meas, n, max_poses, V, split=0, r=3, seed=0, affine=0):
    dh = meas.d + 1
    G = _pose_graph(meas.p1, meas.p2, n)
    plan = emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, max_poses, V, split, affine)
    emu.check_tables(plan)
    A = (pgo.connection_laplacian(meas, n) + 0.1 * sp.identity(dh * n)).tocsc()
    M, Cc, SigInv = emu.dense_blocks(A, plan)
    bufs = emu.fill_stage_buffers(plan, M, Cc, SigInv)
    rv = np.random.default_rng(seed).standard_normal((r, dh * n))
    z = emu.apply(plan, bufs, rv)
    ref = spl.splu(A).solve(rv.T).T
    return plan, np.linalg.norm(z - ref) / np.linalg.norm(ref)


@pytest.mark.parametrize("name,max_poses,V,split", [
    ("tinyGrid3D", 3, 4, 0),          # a handful of poses per domain
    ("smallGrid3D", 20, 7, 0),        # several strips per virtual CTA
    ("smallGrid3D", 0, 148, 0),       # the library's domain size, one B200 worth of CTAs
    ("smallGrid3D", 10, 16, 3),       # forced inner split of the Schur strips
    ("smallGrid3D", 200, 16, 0),      # one domain, no separator: strips longer than one wave of stages
    ("smallGrid3D", 60, 16, 0),       # two large domains: multi-wave interior strips and a separator
])
def test_three_phase_replay_is_exact(datasets, name, max_poses, V, split):
    meas, n, _ = datasets(name)
    plan, err = _replay(meas, n, max_poses, V, split)
    assert err <= 1e-11, err
    if split:
        assert plan["nsplit3"] == min(split, (plan["pcols"] - plan["sep_col0"]) // emu.STAGE_K)


def test_three_phase_replay_sphere2500(datasets):
    """BASELINE configs[1]: 41 domains, 430 separator poses; also the sizing DESIGN.md quotes."""
    meas, n, _ = datasets("sphere2500")
    plan, err = _replay(meas, n, 0, 148, r=2)
    assert err <= 1e-11, err
    assert plan["K"] >= 30 and plan["nS"] <= 0.25 * n
    per_cta = [np.diff(plan["cta" + ph]).max() for ph in ("1", "3", "5")]
    assert max(per_cta) <= 3                               # about one pipeline fill of strips per CTA
    mb = [plan["stages" + ph] * emu.STAGE_K * emu.COLS * 8 / 1e6 for ph in ("1", "3", "5")]
    assert sum(mb) < 80                                    # L2 resident on B200 (126 MB)


def test_domain_affine_balancing(datasets):
    """Strips that read the same input slice share virtual CTAs (the kernel stages the slice once per
    CTA); the replay stays exact and no CTA is worse off than two strips of the longest kind."""
    meas, n, _ = datasets("smallGrid3D")
    plan, err = _replay(meas, n, 12, 24, affine=1)
    assert err <= 1e-11, err
    for ph in ("1", "5"):
        cta, st = plan["cta" + ph], plan["strips" + ph]
        assert len(cta) == plan["V"] + 1
        for v in range(plan["V"]):
            mine = st[cta[v]:cta[v + 1]]
            assert len({(int(s[1]), int(s[2])) for s in mine}) <= 1      # one (kc0, nchunks) per CTA
    meas, n, _ = datasets("sphere2500")
    G = _pose_graph(meas.p1, meas.p2, n)
    pa = emu.fetch_plan(_fn(), n, G.indptr, G.indices, meas.d + 1, 0, 148, 0, 1)
    emu.check_tables(pa)
    st, cta = pa["strips1"], pa["cta1"]
    per_cta = [int(st[cta[v]:cta[v + 1], 2].sum()) for v in range(148)]
    assert max(per_cta) <= 2 * int(st[:, 2].max())


def test_isolated_domain_keeps_empty_strips(datasets):
    """A component small enough to be one domain has no separator neighbour: its phase-5 strips are empty
    (0 chunks) but present, so that the phase still completes its poses (w = 0, i.e. z = y)."""
    meas, n, _ = datasets("smallGrid3D")
    ida = np.arange(0, 100)
    idb = np.arange(105, 125)                               # second component, 20 poses
    new = -np.ones(n, dtype=np.int64)
    new[ida] = np.arange(100)
    new[idb] = 100 + np.arange(20)
    ina, inb = (meas.p1 < 100) & (meas.p2 < 100), (meas.p1 >= 105) & (meas.p2 >= 105)
    keep = ina | inb
    sub = pgo.make_measurements(meas.d, new[meas.p1[keep]], new[meas.p2[keep]], meas.R[keep], meas.t[keep],
                                meas.kappa[keep], meas.tau[keep])
    plan, err = _replay(sub, 120, 20, 6)
    assert err <= 1e-11, err
    grp = set(plan["group"][100:].tolist())
    assert len(grp) == 1 and -1 not in grp                  # the small component is one interior domain
    k = grp.pop()
    assert plan["t_m"][k] == 0
    mine = plan["strips5"][plan["strips5"][:, 6] == k]
    assert len(mine) == plan["dom_pad"][k] // emu.COLS and np.all(mine[:, 2] == 0)


def test_three_phase_plan_2d(datasets):
    """d = 2 (three scalars per pose: tiles straddle the 32- and 64-wide blocks)."""
    meas, n, _ = datasets("city10000")
    keep = (meas.p1 < 400) & (meas.p2 < 400)
    sub = pgo.make_measurements(meas.d, meas.p1[keep], meas.p2[keep], meas.R[keep], meas.t[keep],
                                meas.kappa[keep], meas.tau[keep])
    plan, err = _replay(sub, 400, 30, 12)
    assert plan["dh"] == 3 and err <= 1e-11, err


def test_three_phase_plan_without_separator():
    """A graph below the domain threshold: one domain, no separator, phases 3 and 5 are empty."""
    n, dh = 12, 4
    G = _pose_graph(np.arange(n - 1), np.arange(1, n), n)
    plan = emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, 0, 8)
    emu.check_tables(plan)
    emu.check_tables(emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, 0, 8, 0, 1))
    assert plan["K"] == 1 and plan["nS"] == 0
    assert len(plan["strips3"]) == 0 and len(plan["strips5"]) == 0 and plan["stages5"] == 0
    assert plan["ycols"] == plan["pcols"] == 64


def test_three_phase_plan_rejects_bad_input():
    fn = _fn()
    ip, lp = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    rowptr = np.array([0, 1, 2], dtype=np.int32)
    colidx = np.array([0, 7], dtype=np.int32)                # column out of range
    need = C.c_int64()
    assert fn(2, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), 4, 0, 8, 0, 0, None, 0, C.byref(need)) == -1
    colidx[1] = 1
    assert fn(2, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), 7, 0, 8, 0, 0, None, 0, C.byref(need)) == -1
    assert fn(2, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), 4, 0, 8, 0, 0, None, 0, C.byref(need)) == 0
    assert need.value > 0


# ---- the index arithmetic the CUDA kernels share with the host (dpgo_b200/csrc/dd_stage.h) --------------
@pytest.fixture(scope="module")
def host_arith(tmp_path_factory):
    """tests/native/three_phase_host.cpp compiled with g++ (seconds): dd_stage.h outside nvcc."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path_factory.mktemp("native") / "libthree_phase_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-fvisibility=hidden", "-fno-gnu-unique",
                           "-Wl,-Bsymbolic",
                           os.path.join(root, "tests", "native", "three_phase_host.cpp"), "-o", so])
    return C.CDLL(so)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_kernel_staging_arithmetic_matches_the_replay(datasets, host_arith):
    """strip_stage_value<1/2/3> (the staging step of phase_strip_gemv in the three-phase form) against the
    numpy replay's src1 / src3 / src5 on every inner index the strips of the plan touch."""
    meas, n, _ = datasets("smallGrid3D")
    dh, R = meas.d + 1, 5
    G = _pose_graph(meas.p1, meas.p2, n)
    plan = emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, 12, 24, 2)
    rng = np.random.default_rng(3)
    icol = plan["icol"].astype(np.int32); gidx = plan["gidx"].astype(np.int32)
    tptr = plan["tptr"].astype(np.int32); tcol = plan["tcol"].astype(np.int32)
    ycols, pcols, ns3, sep0 = plan["ycols"], plan["pcols"], plan["nsplit3"], plan["sep_col0"]
    # arrays in the kernels' layout: column-major R x cols, i.e. [col][q]
    r = rng.standard_normal((dh * n, R))
    y = rng.standard_normal((ycols, R))
    zs = rng.standard_normal((ns3, pcols, R))
    assert ns3 >= 2 and len(tcol) > 0
    fn = host_arith.tp_stage_values
    fn.restype = C.c_int

    def call(src, idx, vec):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.zeros((len(idx), R))
        rc = fn(src, R, len(idx), _ip(idx), _dp(vec), _ip(icol), _ip(gidx), _dp(y), _ip(tptr), _ip(tcol), sep0,
                ns3, C.c_int64(pcols * R), _dp(out))
        assert rc == 0
        return out

    def touched(ph):
        st = plan["strips" + ph]
        return np.unique(np.concatenate([np.arange(32 * s[1], 32 * (s[1] + s[2])) for s in st]))

    # phase 1: gathered through icol
    idx = touched("1")
    ref = np.where(icol[idx][:, None] >= 0, r[np.maximum(icol[idx], 0)], 0.0)
    assert np.array_equal(call(1, idx, r), ref)
    # phase 3: r_S minus the scattered coupling terms, in CSR order
    idx = touched("3")
    assert idx.min() >= sep0 and idx.max() < pcols
    ref = np.where(icol[idx][:, None] >= 0, r[np.maximum(icol[idx], 0)], 0.0)
    for a, col in enumerate(idx):
        for e in range(tptr[col - sep0], tptr[col - sep0 + 1]):
            ref[a] -= y[tcol[e]]
    assert np.array_equal(call(2, idx, r), ref)
    # phase 5: partial slots of z_S summed in slot order through the gather list
    idx = touched("5")
    assert idx.max() < len(gidx)
    ref = np.zeros((len(idx), R))
    for a, i in enumerate(idx):
        if gidx[i] >= 0:
            for sl in range(ns3):
                ref[a] += zs[sl, gidx[i]]
    assert np.array_equal(call(3, idx, zs), ref)


def test_kernel_layout_arithmetic_matches_the_replay(datasets, host_arith):
    """layout_rect_value (k_dd_layout_rect) against the stage buffers the numpy replay fills, for the
    coupling strips of both phases, from C_k given with all separator columns + a column map (as the plan
    defines it) and from C_k given on the columns of S_k only (as the device set-up calls it)."""
    meas, n, _ = datasets("smallGrid3D")
    dh = meas.d + 1
    G = _pose_graph(meas.p1, meas.p2, n)
    plan = emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, 12, 24)
    A = (pgo.connection_laplacian(meas, n) + 0.1 * sp.identity(dh * n)).tocsc()
    M, Cc, SigInv = emu.dense_blocks(A, plan)
    bufs = emu.fill_stage_buffers(plan, M, Cc, SigInv)
    fn = host_arith.tp_layout_rect
    fn.restype = C.c_int
    mS = plan["nS"] * dh
    checked = 0
    for ph, kind, form in (("1", 1, 0), ("5", 3, 1)):
        st = plan["strips" + ph]
        for k in range(plan["K"]):
            mine = st[(st[:, 5] == kind) & (st[:, 6] == k)]
            if not len(mine):
                continue
            mine = mine[np.argsort(mine[:, 7])]                      # by output block
            nob, nch, base = len(mine), int(mine[0, 2]), int(mine[0, 4])
            assert np.array_equal(mine[:, 4], base + nch * np.arange(nob))   # [ob][c] runs, as the set-up assumes
            m, tm = int(plan["dom_m"][k]), int(plan["t_m"][k])
            sk = plan["sk"][plan["sk_ptr"][k]:plan["sk_ptr"][k + 1]]
            cmap = (sk[:, None] * dh + np.arange(dh)).ravel().astype(np.int32)
            want = bufs[ph][base:base + nob * nch].ravel()
            Ck = np.asfortranarray(Cc[k])                            # m x tm, the columns of S_k only
            got = np.zeros(nob * nch * 32 * 64)
            assert fn(_dp(Ck), m, m, None, tm, form, nob, nch, _dp(got)) == 0
            assert np.array_equal(got, want)
            full = np.zeros((m, mS), order="F")                      # all separator columns + the column map
            full[:, cmap] = Cc[k]
            got2 = np.zeros_like(got)
            assert fn(_dp(full), m, m, _ip(cmap), tm, form, nob, nch, _dp(got2)) == 0
            assert np.array_equal(got2, want)
            checked += 1
    assert checked >= 4


# ---- the device functions themselves, executed on the host (tests/native/cuda_emu.h) -----------------------
@pytest.fixture(scope="module")
def device_emu_so(tmp_path_factory):
    """kernels.cuh compiled with g++ against a minimal CUDA execution model (one CTA, 256 real threads)."""
    import subprocess
    so = str(tmp_path_factory.mktemp("native") / "libthree_phase_device_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                           "-fvisibility=hidden", "-fno-gnu-unique", "-Wl,-Bsymbolic",
                           os.path.join(ROOT, "tests", "native", "three_phase_device_emu.cpp"), "-o", so])
    return so


def _load_device_emu(so):
    lib = C.CDLL(so)
    lib.tp_apply_device_emu.restype = C.c_int
    return lib


def _in_fresh_interpreter(case, *args):
    """The emulation runs 256 OS threads that meet at barriers thousands of times; it must not share a process
    with whatever thread pools earlier tests have started (it crawls there), so every case gets its own."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.abspath(__file__), case, json.dumps(args)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


STRIP_DT = np.dtype([("cb", np.int32), ("kc0", np.int32), ("nchunks", np.int32), ("slot", np.int32),
                     ("data_off", np.int64)])


def _device_apply(lib, plan, bufs, R, d, Y, rvec, fused, prefetch):
    """One application through phase_strip_gemv<SRC 1 / 2 / 3> (+ fused finish) / phase_dd_finish(_sep)."""
    assert STRIP_DT.itemsize == 24
    keep = []                                               # buffers must outlive the call

    def strips(ph):
        st = plan["strips" + ph]
        a = np.zeros(max(len(st), 1), dtype=STRIP_DT)
        for k, name in enumerate(("cb", "kc0", "nchunks", "slot", "data_off")):
            a[name][:len(st)] = st[:, k]
        cta = np.ascontiguousarray(plan["cta" + ph], dtype=np.int32)
        chunks = np.array([st[cta[v]:cta[v + 1], 2].sum() for v in range(plan["V"])], dtype=np.int32)
        M = np.ascontiguousarray(bufs[ph].reshape(-1)) if bufs[ph].size else np.zeros(1)
        keep.extend([a, cta, chunks, M])
        return [_dp(M), a.ctypes.data_as(C.c_void_p), _ip(cta), _ip(chunks)]

    i32 = lambda name: np.ascontiguousarray(plan[name], dtype=np.int32)
    gidx, icol, tptr, tcol, pcol, srow = (i32(k) for k in ("gidx", "icol", "tptr", "tcol", "pcol", "srow"))
    if not len(gidx):
        gidx = np.zeros(1, dtype=np.int32)
    if not len(tcol):
        tcol = np.zeros(1, dtype=np.int32)
    if not len(srow):
        srow = np.zeros(1, dtype=np.int32)
    N = rvec.shape[1]
    col = lambda A: np.ascontiguousarray(A.T).reshape(-1)   # column-major R x N, as on the device
    y = np.zeros(R * plan["ycols"]); zs = np.zeros(plan["nsplit3"] * R * plan["pcols"]); w = np.zeros(R * plan["pcols"])
    Yc, rc = col(Y), col(rvec)
    z = np.full(R * N, np.nan); neg = np.full(R * N, np.nan); zr = np.zeros(1)
    rc_ = lib.tp_apply_device_emu(R, d, plan["n"], plan["V"], plan["nS"], plan["nsplit3"], plan["sep_col0"],
                                  plan["pcols"], prefetch, *strips("1"), *strips("3"), *strips("5"),
                                  _ip(gidx), _ip(icol), _ip(tptr), _ip(tcol), _ip(pcol), _ip(srow),
                                  _dp(y), _dp(zs), _dp(w), _dp(Yc), _dp(rc), _dp(z), _dp(neg), _dp(zr), int(fused))
    assert rc_ == 0
    return z.reshape(N, R).T, neg.reshape(N, R).T, float(zr[0])


@pytest.mark.parametrize("name,max_poses,V,R,fused,prefetch,split", [
    ("smallGrid3D", 12, 24, 5, 1, 1, -1),     # split -1: domain-affine placement (strips of a domain stage once)
    ("smallGrid3D", 12, 5, 5, 0, 1, 2),       # mode 3: separate finish
    ("smallGrid3D", 12, 5, 5, 1, 1, 2),       # mode 4: finish in the epilogue of the last strip phase
    ("smallGrid3D", 12, 3, 3, 1, 0, 0),       # r = d, no prefetch before the barriers
    ("smallGrid3D", 60, 4, 5, 1, 1, 0),       # domains longer than one wave of stages
    ("smallGrid3D", 200, 2, 5, 1, 1, 0),      # one domain, no separator
    ("tinyGrid3D", 3, 2, 5, 1, 1, 0),
])
def test_device_functions_on_the_host(device_emu_so, name, max_poses, V, R, fused, prefetch, split):
    """z = Proj_Y(r (Q + 0.1 I)^-1), <z, r> and -z from the CUDA device functions run on the host, against
    the oracle's exact preconditioner (src/QuadraticProblem.cpp:56-69)."""
    _in_fresh_interpreter("3d", device_emu_so, name, max_poses, V, R, fused, prefetch, split)


def _device_case(so, name, max_poses, V, R, fused, prefetch, split):
    from conftest import load_dataset
    device_emu = _load_device_emu(so)
    meas, n, _ = load_dataset(name)
    d, dh = meas.d, meas.d + 1
    G = _pose_graph(meas.p1, meas.p2, n)
    plan = emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, max_poses, V, max(split, 0), 1 if split < 0 else 0)
    if split < 0:   # the reuse path must actually occur: the one emulated CTA walks the virtual CTAs in order,
        st = plan["strips1"]                     # so consecutive strips of one domain share their staged slice
        assert sum(tuple(st[i, 1:3]) == tuple(st[i + 1, 1:3]) for i in range(len(st) - 1)) >= 3
    Q = pgo.connection_laplacian(meas, n)
    A = (Q + 0.1 * sp.identity(dh * n)).tocsc()
    bufs = emu.fill_stage_buffers(plan, *emu.dense_blocks(A, plan))
    rng = np.random.default_rng(7)
    Y = pgo.manifold_project(rng.standard_normal((R, dh * n)), d)
    rvec = pgo.tangent_project(Y, rng.standard_normal((R, dh * n)), d)
    ref = pgo.QuadraticProblem(Q, np.zeros((R, dh * n)), d).precondition(Y, rvec)
    z, neg, zr = _device_apply(device_emu, plan, bufs, R, d, Y, rvec, fused, prefetch)
    assert np.linalg.norm(z - ref) <= 1e-10 * np.linalg.norm(ref)
    assert np.array_equal(neg, -z)
    assert abs(zr - float(np.sum(ref * rvec))) <= 1e-10 * abs(float(np.sum(ref * rvec)))


def test_device_functions_on_the_host_2d(device_emu_so):
    """d = 2 (pose tiles of 3 columns straddle the 64-wide strips: separate finish only)."""
    _in_fresh_interpreter("2d", device_emu_so)


def _device_case_2d(so):
    from conftest import load_dataset
    device_emu = _load_device_emu(so)
    meas, n, _ = load_dataset("city10000")
    keep = (meas.p1 < 300) & (meas.p2 < 300)
    sub = pgo.make_measurements(meas.d, meas.p1[keep], meas.p2[keep], meas.R[keep], meas.t[keep],
                                meas.kappa[keep], meas.tau[keep])
    n, d, dh, R = 300, 2, 3, 3
    G = _pose_graph(sub.p1, sub.p2, n)
    plan = emu.fetch_plan(_fn(), n, G.indptr, G.indices, dh, 30, 6)
    Q = pgo.connection_laplacian(sub, n)
    A = (Q + 0.1 * sp.identity(dh * n)).tocsc()
    bufs = emu.fill_stage_buffers(plan, *emu.dense_blocks(A, plan))
    rng = np.random.default_rng(9)
    Y = pgo.manifold_project(rng.standard_normal((R, dh * n)), d)
    rvec = pgo.tangent_project(Y, rng.standard_normal((R, dh * n)), d)
    ref = pgo.QuadraticProblem(Q, np.zeros((R, dh * n)), d).precondition(Y, rvec)
    z, neg, zr = _device_apply(device_emu, plan, bufs, R, d, Y, rvec, 0, 1)
    assert np.linalg.norm(z - ref) <= 1e-10 * np.linalg.norm(ref)
    assert np.array_equal(neg, -z)


def test_rgd_scaled_retraction_on_the_host(device_emu_so):
    """k_retract_scaled (gradient scale applied inside the QF retraction, the RGD step of
    QuadraticOptimizer::gradientDescent, ref: src/QuadraticOptimizer.cpp:133-134) gives the same bits as
    scaling into a separate array and retracting, and a point on the manifold."""
    _in_fresh_interpreter("retract", device_emu_so)


def _retract_case(so):
    lib = _load_device_emu(so)
    lib.emu_retract_scaled.restype = C.c_int
    for R, d, n in ((5, 3, 700), (3, 2, 333)):
        rng = np.random.default_rng(R)
        dh = d + 1
        X = pgo.manifold_project(rng.standard_normal((R, dh * n)), d)
        Dir = rng.standard_normal((R, dh * n))
        col = lambda A: np.ascontiguousarray(A.T).reshape(-1)
        Xc, Dc = col(X), col(Dir)
        a, b = np.zeros(R * dh * n), np.zeros(R * dh * n)
        assert lib.emu_retract_scaled(R, d, n, _dp(Xc), _dp(Dc), C.c_double(-1e-3), _dp(a), _dp(b)) == 0
        assert np.array_equal(a, b)
        out = a.reshape(dh * n, R).T
        ref = pgo.retract_qf(X, -1e-3 * Dir, d)
        assert np.linalg.norm(out - ref) <= 1e-12 * np.linalg.norm(ref)
        T = out.T.reshape(n, dh, R)[:, :d, :]                     # Stiefel blocks: Y^T Y = I
        assert np.abs(np.einsum("nar,nbr->nab", T, T) - np.eye(d)).max() < 1e-12


if __name__ == "__main__":
    import json
    {"3d": _device_case, "2d": _device_case_2d, "retract": _retract_case}[sys.argv[1]](*json.loads(sys.argv[2]))
