"""dpgo_b200 -- B200-native RBCD local-solve hot path of mit-acl/dpgo.

The product is the native library `libdpgo_b200.so` (C-ABI in include/dpgo_b200.h, CUDA kernels
in dpgo_b200/csrc) and the C++ drop-in host API under dpgo_b200/host.  The Python modules here
only bind the C-ABI for tests and bench.py.  There is no CPU fallback anywhere in this package.
"""
from .api import (DeviceProblem, chordal_initialization, default_params, problem_from_measurements,  # noqa: F401
                  SLOT_X, SLOT_Y, SLOT_V, SLOT_XPREV)
from ._lib import DpgoError  # noqa: F401
