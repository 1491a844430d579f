"""CPU check of the in-tree dense linear algebra behind the preconditioner set-up
(dpgo_b200/csrc/dense_la.cuh + dense_la_seq.h; ref: src/PoseGraph.cpp:598-613).  The same device functions and
launch sequences nvcc compiles are built with g++ against tests/native/cuda_emu.h and compared with numpy.
A regression test for rounds without a GPU; the device run is tests/test_gpu_a_parity.py (operator vs oracle)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "dense_la_emu.cpp")
LIB = os.path.join(ROOT, "tests", "native", "libdense_la_emu.so")
DEPS = [SRC, os.path.join(ROOT, "tests", "native", "cuda_emu.h")] + [
    os.path.join(ROOT, "dpgo_b200", "csrc", f) for f in ("dense_la.cuh", "dense_la_seq.h")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               "-fvisibility=hidden", "-Wl,-Bsymbolic", SRC, "-o", LIB])
    L = ctypes.CDLL(LIB)
    dp = ctypes.POINTER(ctypes.c_double)
    L.dla_emu_spd_inverse.argtypes = [dp, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]
    L.dla_emu_gemm.argtypes = [dp, dp, dp] + [ctypes.c_int] * 10 + [ctypes.c_double] * 2
    return L


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _spd(rng, n):
    M = rng.standard_normal((n, n))
    return M @ M.T / n + 0.1 * np.eye(n)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_tile_gemm_all_transposes(lib, ta, tb):
    rng = np.random.default_rng(ta * 2 + tb)
    M, N, K = 70, 133, 37
    opA, opB = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    A = np.asfortranarray(opA.T if ta else opA)
    B = np.asfortranarray(opB.T if tb else opB)
    C0 = rng.standard_normal((M + 3, N))
    C = np.asfortranarray(C0.copy())
    lib.dla_emu_gemm(_ptr(A), _ptr(B), _ptr(C), M, N, K, A.shape[0], B.shape[0], C.shape[0], ta, tb, 0, 0, -0.5, 2.0)
    ref = C0.copy()
    ref[:M] = -0.5 * opA @ opB + 2.0 * C0[:M]
    assert np.allclose(C, ref, rtol=1e-13, atol=1e-13)


def test_tile_gemm_triangular_k_ranges(lib):
    """The k-range modes skip exactly the tiles a triangular operand makes zero (garbage above the diagonal
    tiles must not be read)."""
    rng = np.random.default_rng(7)
    n = 150
    Lo = np.tril(rng.standard_normal((n, n)))
    tile = np.arange(n) // 64
    junk = np.asfortranarray(np.where(tile[None, :] > tile[:, None], np.nan, Lo))   # NaN in every tile above the diagonal tiles
    X = np.asfortranarray(rng.standard_normal((n, n)))
    C = np.asfortranarray(np.zeros((n, n)))
    lib.dla_emu_gemm(_ptr(X), _ptr(junk), _ptr(C), n, n, n, n, n, n, 0, 0, 0, 1, 1.0, 0.0)     # X * L
    assert np.allclose(C, X @ Lo, rtol=1e-12, atol=1e-12)
    lib.dla_emu_gemm(_ptr(junk), _ptr(X), _ptr(C), n, n, n, n, n, n, 0, 0, 0, 2, 1.0, 0.0)     # L * X
    assert np.allclose(C, Lo @ X, rtol=1e-12, atol=1e-12)
    C[:] = 0
    lib.dla_emu_gemm(_ptr(junk), _ptr(junk), _ptr(C), n, n, n, n, n, n, 1, 0, 1, 3, 1.0, 0.0)  # L^T L, lower tiles
    ref = Lo.T @ Lo
    mask = np.add.outer(np.arange(n) // 64, -(np.arange(n) // 64)) >= 0
    assert np.allclose(C[mask], ref[mask], rtol=1e-12, atol=1e-12)
    assert np.all(C[~mask] == 0)


@pytest.mark.parametrize("sizes", [[1], [5, 64, 65], [130, 200, 17, 0, 321]])
def test_batched_spd_inverse(lib, sizes):
    rng = np.random.default_rng(len(sizes))
    mats = [_spd(rng, n) for n in sizes]
    buf = np.concatenate([np.asfortranarray(np.tril(M)).ravel(order="F") for M in mats] + [np.zeros(1)])
    n = (ctypes.c_int * len(sizes))(*sizes)
    assert lib.dla_emu_spd_inverse(_ptr(buf), n, len(sizes), 0) == 0
    off = 0
    for M, k in zip(mats, sizes):
        got = buf[off:off + k * k].reshape((k, k), order="F")
        off += k * k
        ref = np.linalg.inv(M)
        assert np.allclose(np.tril(got), np.tril(ref), rtol=1e-9, atol=1e-11), k


def test_spd_inverse_symmetrized_and_failure_code(lib):
    rng = np.random.default_rng(3)
    M = _spd(rng, 100)
    buf = np.asfortranarray(np.tril(M)).ravel(order="F").copy()
    n = (ctypes.c_int * 1)(100)
    assert lib.dla_emu_spd_inverse(_ptr(buf), n, 1, 1) == 0
    got = buf.reshape((100, 100), order="F")
    assert np.allclose(got, np.linalg.inv(M), rtol=1e-9, atol=1e-11)
    bad = np.eye(70)
    bad[66, 66] = -1.0
    both = np.concatenate([np.asfortranarray(np.tril(M)).ravel(order="F"), bad.ravel(order="F")])
    n2 = (ctypes.c_int * 2)(100, 70)
    assert lib.dla_emu_spd_inverse(_ptr(both), n2, 2, 0) == 2      # 1 + index of the indefinite matrix
