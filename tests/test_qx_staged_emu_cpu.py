"""CPU check of the shared-memory staged Q*X kernel (dpgo_b200/csrc/qx_staged.cuh; ref:
src/QuadraticProblem.cpp:29-54): the device function nvcc compiles is built with g++ against
tests/native/cuda_emu.h and compared with the oracle's connection Laplacian.  A regression test for rounds
without a GPU; the device run is tests/test_gpu_a_parity.py::test_qx_staged_parity."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pgo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "qx_staged_emu.cpp")
LIB = os.path.join(ROOT, "tests", "native", "libqx_staged_emu.so")
DEPS = [SRC, os.path.join(ROOT, "tests", "native", "cuda_emu.h")] + [
    os.path.join(ROOT, "dpgo_b200", "csrc", f) for f in ("qx_staged.cuh", "kernels.cuh")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               "-fvisibility=hidden", "-Wl,-Bsymbolic", "-I" + os.path.join(ROOT, "dpgo_b200", "csrc"),
                               SRC, "-o", LIB])
    return ctypes.CDLL(LIB)


def bsr_of(Qm, n, dh):
    """block-CSR with row-major blocks, as dpgo_finalize lays Q out"""
    B = sp.bsr_matrix(sp.csr_matrix(Qm), blocksize=(dh, dh))
    B.sort_indices()
    return (np.ascontiguousarray(B.indptr, dtype=np.int32), np.ascontiguousarray(B.indices, dtype=np.int32),
            np.ascontiguousarray(B.data, dtype=np.float64))


def _run(lib, r, d, rowptr, colidx, blocks, X, G, n, ctas):
    out = np.full((r, (d + 1) * n), np.nan, order="F")
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    rc = lib.qx_staged_emu(r, d, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), blocks.ctypes.data_as(dp),
                           X.ctypes.data_as(dp), None if G is None else G.ctypes.data_as(dp),
                           out.ctypes.data_as(dp), n, ctas)
    assert rc == 0
    return out


@pytest.mark.parametrize("name,r,ctas", [("tinyGrid3D", 3, 1), ("smallGrid3D", 5, 1), ("smallGrid3D", 5, 3)])
def test_staged_qx_matches_the_oracle(lib, datasets, name, r, ctas):
    meas, n, _ = datasets(name)
    d = meas.d
    Qm = pgo.connection_laplacian(meas, n)
    rowptr, colidx, blocks = bsr_of(Qm, n, d + 1)
    rng = np.random.default_rng(3)
    X = np.asfortranarray(rng.standard_normal((r, (d + 1) * n)))
    G = np.asfortranarray(rng.standard_normal((r, (d + 1) * n)))
    ref = X @ Qm
    got = _run(lib, r, d, rowptr, colidx, blocks, X, None, n, ctas)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    got = _run(lib, r, d, rowptr, colidx, blocks, X, G, n, ctas)
    assert np.allclose(got, ref + G, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("d,r", [(2, 3), (2, 4), (3, 5)])
def test_staged_qx_long_rows_and_few_rows(lib, d, r):
    """rows longer than one step (hub poses with > 32/(d+1) blocks), fewer rows than warps, 2-D tiles (8-byte pieces)"""
    rng = np.random.default_rng(d * 10 + r)
    n, dh = 37, d + 1
    A = sp.random(n, n, density=0.15, random_state=5, format="lil")
    A[0, :] = 1.0                                   # a hub: 37 blocks in row 0 (and column 0)
    A = ((A + A.T) != 0).astype(float) + sp.eye(n)
    dense = np.kron(A.toarray() != 0, np.ones((dh, dh))) * rng.standard_normal((n * dh, n * dh))
    dense = dense + dense.T
    rowptr, colidx, blocks = bsr_of(dense, n, dh)
    assert (np.diff(rowptr) > 32 // dh).any()
    X = np.asfortranarray(rng.standard_normal((r, dh * n)))
    for ctas in (1, 7):
        got = _run(lib, r, d, rowptr, colidx, blocks, X, None, n, ctas)
        assert np.allclose(got, X @ dense, rtol=1e-12, atol=1e-11)
