#!/usr/bin/env python
"""Driver for ncu captures of one preconditioner application: python tools/ncu_precon_apply.py <dataset> <r> <mode>
applies the operator 20 times through the stand-alone kernels (the same device functions the fused solver inlines)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from bench import load_fixture  # noqa: E402

name, r, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
z, d, n = load_fixture(name)
gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r, precon_mode=mode)
print("apply_us", gp.time_precon(20, False))
gp.close()
