"""CPU check of the Q*X device functions (dpgo_b200/csrc/kernels.cuh; ref: src/QuadraticProblem.cpp:29-54): the
lane-group kernel, its two-blocks-per-step form and the form that stages the gathered pose tiles in shared memory are
built with g++ against tests/native/cuda_emu.h and compared with the oracle's connection Laplacian -- and with each
other bit for bit (same sums in the same order).  A regression test for rounds without a GPU; the device run is
tests/test_gpu_a_parity.py::test_qx_variants_parity."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pgo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "qx_tiles_emu.cpp")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("native") / "libqx_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                           "-fvisibility=hidden", "-Wl,-Bsymbolic", SRC, "-o", so])
    return ctypes.CDLL(so)


def bsr_of(Qm, dh):
    """block-CSR with row-major blocks, as dpgo_finalize lays Q out"""
    B = sp.bsr_matrix(sp.csr_matrix(Qm), blocksize=(dh, dh))
    B.sort_indices()
    return (np.ascontiguousarray(B.indptr, dtype=np.int32), np.ascontiguousarray(B.indices, dtype=np.int32),
            np.ascontiguousarray(B.data, dtype=np.float64))


def _run(lib, variant, r, d, rowptr, colidx, blocks, X, G, n, ctas):
    out = np.full((r, (d + 1) * n), np.nan, order="F")
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    rc = lib.qx_emu(variant, r, d, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), blocks.ctypes.data_as(dp),
                    X.ctypes.data_as(dp), None if G is None else G.ctypes.data_as(dp), out.ctypes.data_as(dp), n, ctas)
    assert rc == 0
    return out


@pytest.mark.parametrize("name,r,ctas", [("tinyGrid3D", 3, 1), ("smallGrid3D", 5, 1), ("smallGrid3D", 5, 3)])
def test_qx_forms_match_the_oracle_and_each_other(lib, datasets, name, r, ctas):
    meas, n, _ = datasets(name)
    d = meas.d
    Qm = pgo.connection_laplacian(meas, n)
    rowptr, colidx, blocks = bsr_of(Qm, d + 1)
    rng = np.random.default_rng(3)
    X = np.asfortranarray(rng.standard_normal((r, (d + 1) * n)))
    G = np.asfortranarray(rng.standard_normal((r, (d + 1) * n)))
    ref = X @ Qm
    base = _run(lib, 0, r, d, rowptr, colidx, blocks, X, None, n, ctas)
    assert np.allclose(base, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    baseG = _run(lib, 0, r, d, rowptr, colidx, blocks, X, G, n, ctas)
    for variant in (2, 3):
        assert np.array_equal(_run(lib, variant, r, d, rowptr, colidx, blocks, X, None, n, ctas), base), variant
        assert np.array_equal(_run(lib, variant, r, d, rowptr, colidx, blocks, X, G, n, ctas), baseG), variant


@pytest.mark.parametrize("d,r", [(2, 3), (2, 4), (3, 5)])
def test_qx_forms_ragged_rows(lib, d, r):
    """rows of very different lengths inside one warp (a hub pose), fewer rows than lane groups, 2-D tiles (8-byte
    pieces when the tile has an odd number of doubles)"""
    rng = np.random.default_rng(d * 10 + r)
    n, dh = 37, d + 1
    A = sp.random(n, n, density=0.15, random_state=5, format="lil")
    A[0, :] = 1.0                                   # a hub: 37 blocks in row 0 (and column 0)
    A = ((A + A.T) != 0).astype(float) + sp.eye(n)
    dense = np.kron(A.toarray() != 0, np.ones((dh, dh))) * rng.standard_normal((n * dh, n * dh))
    dense = dense + dense.T
    rowptr, colidx, blocks = bsr_of(dense, dh)
    X = np.asfortranarray(rng.standard_normal((r, dh * n)))
    for ctas in (1, 2):
        base = _run(lib, 0, r, d, rowptr, colidx, blocks, X, None, n, ctas)
        assert np.allclose(base, X @ dense, rtol=1e-12, atol=1e-11)
        for variant in (2, 3):
            assert np.array_equal(_run(lib, variant, r, d, rowptr, colidx, blocks, X, None, n, ctas), base), variant


@pytest.mark.parametrize("d,r,n", [(3, 5, 300), (2, 3, 77), (2, 4, 256)])
def test_staged_pose_kernels_match_the_per_pose_functions(lib, d, r, n):
    """The stand-alone QF retraction / polar projection / rounding kernels stage the 32 tiles of a warp step through
    shared memory (pose_staged: coalesced global traffic at scale) and run the same per-pose function on the staged
    values: bit-identical to the function applied tile by tile, including a ragged last warp step (n not a multiple
    of 32) and odd tile sizes; and equal to the oracle (ref: ProductManifold::Retraction, LiftedSEManifold::project
    src/manifold/LiftedSEManifold.cpp:34-45, projectToRotationGroup src/DPGO_utils.cpp:464-478)."""
    dh = d + 1
    rng = np.random.default_rng(100 * d + r)
    X = pgo.manifold_project(rng.standard_normal((r, dh * n)), d)
    B = X + 0.05 * rng.standard_normal(X.shape)
    Cm = X + 0.05 * rng.standard_normal(X.shape)
    A_, B_, C_ = (np.asfortranarray(a) for a in (X, B, Cm))
    dp = ctypes.POINTER(ctypes.c_double)
    outs = {}
    for op, cols in ((0, r), (1, r), (2, d)):
        for staged in (0, 1):
            out = np.full((cols, dh * n), np.nan, order="F")
            rc = lib.pose_op_emu(op, staged, r, d, A_.ctypes.data_as(dp), B_.ctypes.data_as(dp), C_.ctypes.data_as(dp),
                                 out.ctypes.data_as(dp), n)
            assert rc == 0
            outs[op, staged] = out
        assert np.array_equal(outs[op, 0], outs[op, 1]), op
    ref = pgo.manifold_project(0.5 * X + 0.3 * B + 0.2 * Cm, d)
    assert np.allclose(outs[1, 1], ref, rtol=0, atol=1e-12)
    ret = outs[0, 1]
    for i in (0, n // 2, n - 1):
        Y = ret[:, i * dh:i * dh + d]
        assert np.allclose(Y.T @ Y, np.eye(d), atol=1e-13)
        assert np.allclose(ret[:, i * dh + d], X[:, i * dh + d] + B[:, i * dh + d], atol=0)
    T = outs[2, 1]
    assert np.allclose(T[:, :dh], np.eye(d, dh), atol=1e-12)          # pose 0 in its own frame
    for i in (1, n - 1):
        Ri = T[:, i * dh:i * dh + d]
        assert np.allclose(Ri.T @ Ri, np.eye(d), atol=1e-12) and np.linalg.det(Ri) > 0
