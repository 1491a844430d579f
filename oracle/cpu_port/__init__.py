"""ctypes binding of the compiled CPU port (oracle/cpu_port/dpgo_cpu.cpp) -- TEST / BASELINE
INFRASTRUCTURE, see oracle/__init__.py.  Single-threaded C++ restatement of the reference's local
solve: block-CSR Q X, exact block sparse Cholesky preconditioner (minimum-degree ordering),
per-pose Stiefel projection / QF retraction, RTR + tCG."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libdpgo_cpu.so")


class CpuResult(C.Structure):
    _fields_ = [("f_init", C.c_double), ("gn_init", C.c_double), ("f_opt", C.c_double), ("gn_opt", C.c_double),
                ("outer", C.c_int), ("inner", C.c_int), ("accepted", C.c_int), ("rejected", C.c_int),
                ("tcg_status", C.c_int), ("pad", C.c_int), ("qx", C.c_long), ("precon", C.c_long)]


def _load():
    if not os.path.exists(_LIB):
        from oracle import build_oracle
        build_oracle.build()
    lib = C.CDLL(_LIB)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.cpu_create.restype = C.c_void_p
    lib.cpu_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp]
    lib.cpu_destroy.argtypes = [C.c_void_p]
    lib.cpu_set_G.argtypes = [C.c_void_p, dp]
    lib.cpu_factorize.argtypes = [C.c_void_p]
    lib.cpu_factor_blocks.restype = C.c_long
    lib.cpu_factor_blocks.argtypes = [C.c_void_p]
    lib.cpu_qx.argtypes = [C.c_void_p, dp, dp]
    lib.cpu_solve.argtypes = [C.c_void_p, dp, dp]
    lib.cpu_optimize.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.POINTER(CpuResult)]
    return lib


_lib = None


class CpuProblem:
    def __init__(self, Q, G, d):
        global _lib
        if _lib is None:
            _lib = _load()
        self.d, self.r = d, G.shape[0]
        dh = d + 1
        self.n = G.shape[1] // dh
        B = sp.bsr_matrix(Q, blocksize=(dh, dh))
        B.sort_indices()
        self._rowptr = np.ascontiguousarray(B.indptr, dtype=np.int32)
        self._colidx = np.ascontiguousarray(B.indices, dtype=np.int32)
        self._blocks = np.ascontiguousarray(B.data, dtype=np.float64)
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        self._h = C.c_void_p(_lib.cpu_create(self.n, d, self.r, len(self._colidx), self._rowptr.ctypes.data_as(ip),
                                             self._colidx.ctypes.data_as(ip), self._blocks.ctypes.data_as(dp)))
        self.set_G(G)

    def _d(self, a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def set_G(self, G):
        G = np.asfortranarray(G, dtype=np.float64)
        _lib.cpu_set_G(self._h, self._d(G))

    def factorize(self):
        if _lib.cpu_factorize(self._h) != 0:
            raise RuntimeError("Q + 0.1 I is not positive definite")
        return _lib.cpu_factor_blocks(self._h)

    def qx(self, X):
        X = np.asfortranarray(X, dtype=np.float64); out = np.empty_like(X, order="F")
        _lib.cpu_qx(self._h, self._d(X), self._d(out))
        return out

    def solve(self, V):
        V = np.asfortranarray(V, dtype=np.float64); out = np.empty_like(V, order="F")
        _lib.cpu_solve(self._h, self._d(V), self._d(out))
        return out

    def optimize(self, X0, gradnorm_tol=1e-2, RTR_iterations=3, RTR_tCG_iterations=50, RTR_initial_radius=100.0):
        X0 = np.asfortranarray(X0, dtype=np.float64); out = np.empty_like(X0, order="F")
        res = CpuResult()
        rc = _lib.cpu_optimize(self._h, self._d(X0), self._d(out), gradnorm_tol, RTR_iterations,
                               RTR_tCG_iterations, RTR_initial_radius, C.byref(res))
        if rc != 0:
            raise RuntimeError("cpu_optimize failed")
        return out, {k: getattr(res, k) for k, _ in CpuResult._fields_}

    def close(self):
        if self._h:
            _lib.cpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
