"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dpgo_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(headers=("dpgo_b200.h", "dpgo_b200_dev.h")):
    found = set()
    for hname in headers:
        txt = open(os.path.join(ROOT, "include", hname)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        found |= set(re.findall(r"\b(dpgo_[a-zA-Z0-9_]+)\s*\(", txt))
    return sorted(found)


def test_library_exports_every_declared_symbol():
    from dpgo_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/dpgo_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)


def test_struct_layout_matches_header():
    from dpgo_b200 import _lib
    p = _lib.RoptParams()
    _lib.lib.dpgo_default_params(C.byref(p))
    # defaults of ROptParameters, include/DPGO/DPGO_types.h:53-61 of the reference
    assert (p.method, p.gradnorm_tol, p.RGD_stepsize, p.RGD_use_preconditioner) == (0, 1e-2, 1e-3, 1)
    assert (p.RTR_iterations, p.RTR_tCG_iterations, p.RTR_initial_radius) == (3, 50, 100.0)
    assert (p.tcg_theta, p.tcg_kappa, p.accept_rho, p.shrink, p.magnify) == (1.0, 0.1, 0.1, 0.25, 2.0)
    assert C.sizeof(_lib.RoptParams) == 88
    assert C.sizeof(_lib.RoptResult) == 232


def test_contract_violations_return_einval():
    from dpgo_b200 import _lib
    h = C.c_void_p()
    assert _lib.lib.dpgo_create(0, 0, 3, 5, None, C.byref(h)) == -1      # n > 0
    assert _lib.lib.dpgo_create(0, 4, 4, 5, None, C.byref(h)) == -1      # d in {2,3}
    assert _lib.lib.dpgo_create(0, 4, 3, 2, None, C.byref(h)) == -1      # r >= d (PoseGraph.cpp:19)
    assert b"contract violated" in _lib.lib.dpgo_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dpgo_b200 import _lib
    h = C.c_void_p()
    rc = _lib.lib.dpgo_create(0, 4, 3, 5, None, C.byref(h))
    assert rc == -2
    assert b"no CPU fallback" in _lib.lib.dpgo_last_error()
