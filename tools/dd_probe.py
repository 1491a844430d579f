"""Development probe (GPU): storage variants and tuning of the exact preconditioner.

Prints, per problem and configuration, the stand-alone apply time, the fused-solver time of one
optimize() and the in-kernel phase clocks (dpgo_ropt_result.phase_ms), as JSON lines.

    python tools/dd_probe.py [--quick | --barrier-ab] > gpurun_out/dd_probe.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from dpgo_b200 import synthetic  # noqa: E402
from bench import lifting_matrix  # noqa: E402


def fixture(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return dict(z), int(z["d"]), int(z["n"]), z["T_chordal"]


def run(tag, z, d, n, T0, r, mode, tuning=None, reps=3, domain_size=None):
    X0 = np.asfortranarray(lifting_matrix(d, r) @ T0)
    t0 = time.time()
    gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r,
                                             precon_mode=mode, precon_tuning=tuning, domain_size=domain_size)
    setup = time.time() - t0
    gp.slot_set(0, X0)
    us = gp.time_precon(20, False)
    by = gp.bytes_precon()
    best = None
    for _ in range(reps):
        _, res = gp.optimize(X0, dpgo_b200.default_params())
        if best is None or res["elapsed_ms"] < best["elapsed_ms"]:
            best = res
    ph = best["phase_ms"]
    print(json.dumps({"problem": tag, "n": n, "d": d, "mode": mode, "tuning": tuning, "domain_size": domain_size,
                      "setup_s": round(setup, 2),
                      "apply_us": round(us, 1), "apply_bytes": by, "optimize_ms": round(best["elapsed_ms"], 3),
                      "outer": best["outer_iters"], "tcg": best["inner_iters"], "n_precon": best["n_precon"],
                      "two_f": 2 * best["f_opt"], "barriers": best["n_barriers"],
                      "phase_ms": [round(v, 3) for v in ph]}), flush=True)
    gp.close()


def barrier_ab():
    """--barrier-ab: the solves whose time is barrier / latency bound, for A/B runs of library builds that differ
    in the grid barrier (DPGO_B200_LIB=...): bench problem in the two-level forms and with the dense inverse,
    one grid3D-agent-sized problem."""
    z, d, n, T0 = fixture("sphere2500")
    for mode in (2, 0):
        run("sphere2500", z, d, n, T0, 5, mode, reps=5)
    g = synthetic.grid3d(10, seed=1)
    for mode in (0, 2):
        run("grid3d_L10", g, 3, 1000, g["T_true"], 5, mode, reps=5)


def strip_tuning():
    """--strip-tuning: inner splits of the interior / Schur strips (balance over the 148 CTAs against partial sums)."""
    z, d, n, T0 = fixture("sphere2500")
    for tuning in [None, (1, 0, -1), (2, 0, -1), (3, 0, -1), (2, 4, -1), (2, 7, -1), (1, 7, -1), (1, 3, -1)]:
        run("sphere2500", z, d, n, T0, 5, 2, tuning, reps=5)
    g = synthetic.grid3d(10, seed=1)
    for tuning in [None, (2, 0, -1), (3, 0, -1)]:
        run("grid3d_L10", g, 3, 1000, g["T_true"], 5, 2, tuning, reps=5)


def main():
    if "--strip-tuning" in sys.argv:
        return strip_tuning()
    if "--barrier-ab" in sys.argv:
        return barrier_ab()
    quick = "--quick" in sys.argv
    z, d, n, T0 = fixture("sphere2500")
    run("sphere2500", z, d, n, T0, 5, 0)
    for tuning in [(0, 0, 0), (0, 0, 1), (1, 4, 1), (1, 8, 1)]:
        run("sphere2500", z, d, n, T0, 5, 2, tuning)
    z, d, n, T0 = fixture("smallGrid3D")
    for mode in (0, 2):
        run("smallGrid3D", z, d, n, T0, 5, mode)
    # one agent of the grid3D / 8 agents workload has 1000 poses: where is the break-even?
    for L in (10, 12):
        g = synthetic.grid3d(L, seed=1)
        T = g["T_true"] if "T_true" in g else g["T0"]
        for mode in (0, 2):
            run(f"grid3d_L{L}", g, 3, L ** 3, T, 5, mode)
    if not quick:
        for name in ("torus3D", "city10000"):
            z, d, n, T0 = fixture(name)
            for mode in (0, 2):
                run(name, z, d, n, T0, 5 if d == 3 else 3, mode, reps=2)


if __name__ == "__main__":
    main()
