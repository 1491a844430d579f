"""GPU check of the three-phase form of the two-level preconditioner (precon_mode 3), run as its
own process by tests/test_gpu_zzz_three_phase.py: operator against the oracle's exact solve and
against mode 2, full solve (host-driven and fused) against the oracle's iteration counts, and the
apply / solve times of modes 2 and 3 side by side.  Prints one JSON line per data set."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_dataset  # noqa: E402
from oracle import pgo  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main(cases):
    import dpgo_b200
    failed = 0
    for name, r, tunings in cases:
        meas, n, z = load_dataset(name)
        d = meas.d
        rng = np.random.default_rng(11)
        X = pgo.manifold_project(rng.standard_normal((r, (d + 1) * n)), d)
        Vt = pgo.tangent_project(X, rng.standard_normal((r, (d + 1) * n)), d)
        op = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((r, (d + 1) * n)), d)
        ref = op.precondition(X, Vt)
        X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
        Xo, ro = pgo.optimize(op, X0)
        g2 = dpgo_b200.problem_from_measurements(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r,
                                                 precon_mode=2)
        z2 = g2.precon(X, Vt)
        X2, r2 = g2.optimize(X0, dpgo_b200.default_params())
        line = {"dataset": name, "r": r, "mode2": {"apply_us": g2.time_precon(20, False), "solve_ms": r2["elapsed_ms"],
                                                   "bytes": g2.bytes_precon()}}
        g2.close()
        # mode 4 = mode 3 with the finish fused into the last strip phase of the fused solver (d = 3)
        for mode, tuning in [(3, t) for t in tunings] + [(4, None)]:
            g3 = dpgo_b200.problem_from_measurements(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r,
                                                     precon_mode=mode, precon_tuning=tuning)
            assert g3.precon_mode() == (mode if d == 3 else 3)
            z3 = g3.precon(X, Vt)
            e_ref, e_2 = rel(z3, ref), rel(z3, z2)
            ok = e_ref < 1e-8 and e_2 < 1e-9
            sols = {}
            for fused in (0, 1):
                X3, r3 = g3.optimize(X0, dpgo_b200.default_params(fused=fused))
                same_iters = (r3["outer_iters"], r3["inner_iters"]) == (ro.outer, ro.inner_total)
                gap = abs(r3["f_opt"] - ro.fOpt) / abs(ro.fOpt)
                ok = ok and same_iters and gap <= 1e-9 and rel(X3, Xo) < 1e-6
                sols["fused" if fused else "host"] = {"iters": [r3["outer_iters"], r3["inner_iters"]], "gap": gap,
                                                      "solve_ms": r3["elapsed_ms"], "barriers": r3["n_barriers"],
                                                      "phase_ms": r3["phase_ms"][:13]}
            line["mode%d_tuning_%s" % (mode, tuning)] = {"ok": bool(ok), "err_vs_oracle": e_ref, "err_vs_mode2": e_2,
                                                    "apply_us": g3.time_precon(20, False), "bytes": g3.bytes_precon(),
                                                    "solves": sols}
            failed += 0 if ok else 1
            g3.close()
        print(json.dumps(line), flush=True)
    return failed


if __name__ == "__main__":
    # tunings (split_interior, split_schur, prefetch): split_interior == 2 = domain-affine strip placement
    quick = [("tinyGrid3D", 3, [None]), ("smallGrid3D", 5, [None, (0, 3, 0), (2, 0, -1)]),
             ("sphere2500", 5, [None, (0, 7, 0), (2, 0, -1)])]
    full = quick + [("city10000", 3, [None]), ("torus3D", 5, [None])]
    sys.exit(1 if main(full if "--full" in sys.argv else quick) else 0)
