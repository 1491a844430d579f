#!/usr/bin/env python
"""Set-up time of the exact preconditioner (ref: src/PoseGraph.cpp:598-613, rebuilt at every GNC weight update,
src/PGOAgent.cpp:1104-1142): in-tree batched Cholesky / inverse / tile GEMM of dense_la.cu.  One JSON line per case:
wall time of dpgo_finalize(build_precon) minus the same call without the preconditioner."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from bench import load_fixture  # noqa: E402


def agent_slice(z, n, A, a):
    """contiguous split as examples/MultiRobotExample.cpp:71-88: private edges of agent a, re-indexed"""
    per = n // A
    lo, hi = a * per, (n if a == A - 1 else (a + 1) * per)
    m = (z["p1"] >= lo) & (z["p1"] < hi) & (z["p2"] >= lo) & (z["p2"] < hi)
    return dict(p1=z["p1"][m] - lo, p2=z["p2"][m] - lo, R=z["R"][m], t=z["t"][m], kappa=z["kappa"][m],
                tau=z["tau"][m]), hi - lo


def main():
    cases = [("sphere2500", 1, 5, None), ("sphere2500", 1, 5, 0), ("city10000", 4, 3, None), ("city10000", 1, 3, None),
             ("grid3D", 8, 5, None), ("grid3D", 8, 5, 0), ("torus3D", 1, 5, None)]
    for name, A, r, mode in cases:
        z, d, n = load_fixture(name)
        e, na = (dict(p1=z["p1"], p2=z["p2"], R=z["R"], t=z["t"], kappa=z["kappa"], tau=z["tau"]), n) if A == 1 \
            else agent_slice(z, n, A, 0)
        times = {}
        for build in (False, True, True, True):
            t0 = time.perf_counter()
            gp = dpgo_b200.problem_from_measurements(e["p1"], e["p2"], e["R"], e["t"], e["kappa"], e["tau"], na, d, r,
                                                     build_precon=build, precon_mode=mode)
            dt = time.perf_counter() - t0
            times.setdefault(build, []).append(dt)
            pm = gp.precon_mode() if build else None
            gp.close()
        print(json.dumps({"dataset": name, "agents": A, "poses": na, "d": d, "r": r, "precon_mode": pm,
                          "no_precon_s": round(min(times[False]), 4),
                          "with_precon_s": [round(t, 4) for t in times[True]],
                          "precon_setup_s": round(min(times[True]) - min(times[False]), 4)}), flush=True)


if __name__ == "__main__":
    main()
