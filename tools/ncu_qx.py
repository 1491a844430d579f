#!/usr/bin/env python
"""Driver for ncu captures of the stand-alone Q*X at roofline scale: python tools/ncu_qx.py <L> <variant> [reps]
builds the synthetic L^3 grid (dpgo_b200/synthetic.py) and launches the chosen kernel variant `reps` times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from dpgo_b200 import synthetic  # noqa: E402

L, variant = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = synthetic.grid3d(L)
gp = dpgo_b200.problem_from_measurements(g["p1"], g["p2"], g["R"], g["t"], g["kappa"], g["tau"], g["n"], 3, 5,
                                         build_precon=False)
gp.slot_set(0, np.random.default_rng(0).standard_normal((5, 4 * g["n"])))
gp.set_qx_variant(variant, 4096 if variant == 1 else 0)
print("us", gp.time_qx(reps, True))
gp.close()
