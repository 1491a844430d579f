#!/usr/bin/env python
"""Roofline-scale measurement of the per-pose kernels (QF retraction, polar projection of the Nesterov updates,
rounding to SE(d)) on synthetic 3-D grids that do not fit in the L2: CUDA events on the launching stream, L2 flushed
clean before every timed launch.  Algorithmic bytes: retraction reads X and eta and writes X+ (3 tile arrays), the
polar form reads three arrays and writes one (4), rounding reads the lifted poses and writes d x (d+1) poses."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from dpgo_b200 import synthetic  # noqa: E402
from bench import lifting_matrix  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 100]
    peak = 6481.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    d, r = 3, 5
    for L in sizes:
        g = synthetic.grid3d(L)
        n = g["n"]
        gp = dpgo_b200.problem_from_measurements(g["p1"], g["p2"], g["R"], g["t"], g["kappa"], g["tau"], n, d, r,
                                                 build_precon=False)
        rng = np.random.default_rng(0)
        X = np.asfortranarray(lifting_matrix(d, r) @ g["T_true"])
        gp.slot_set(0, X)
        gp.slot_set(1, X + 0.05 * rng.standard_normal(X.shape))
        gp.slot_set(2, X + 0.05 * rng.standard_normal(X.shape))
        tile = r * (d + 1) * n * 8.0
        rec = dict(L=L, n=n, peak_gbs=peak, ops=[])
        for op, name, nbytes in ((0, "retract", 3 * tile), (1, "polar", 4 * tile), (2, "round", tile * (r + d) / r)):
            warm, cold = gp.time_pose_op(op, 20, False), gp.time_pose_op(op, 10, True)
            rec["ops"].append(dict(op=name, bytes=nbytes, back_to_back_us=warm, flushed_us=cold,
                                   flushed_gbs=nbytes / cold / 1e3, frac_of_measured_peak=nbytes / cold / 1e3 / peak))
        print(json.dumps(rec), flush=True)
        gp.close()


if __name__ == "__main__":
    main()
