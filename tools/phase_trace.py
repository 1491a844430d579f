"""Per-CTA phase times of the fused solver (measurement build of the library):

    python dpgo_b200/build.py --trace && python tools/phase_trace.py [dataset] [r] [precon_mode ...]
    python dpgo_b200/build.py            # back to the normal build afterwards

For every phase id of dpgo_ropt_result.phase_ms: the time CTA 0 saw (work + waiting at the barrier,
summed over the solve), and the distribution over the CTAs of the time each one actually worked before
reaching the barrier.  total - max(busy) is what the barriers themselves cost; max(busy) against
median(busy) is the imbalance of the phase.  One JSON line per preconditioner variant."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PHASES = ["cost_grad", "precon_stream", "precon_finish", "hessvec", "tcg_update", "tcg_direction", "retract_copy",
          "unused", "dd_phase1", "dd_sep_rhs", "dd_schur", "dd_back_rhs", "dd_last_strips"]


def trace(prob):
    from dpgo_b200._lib import lib, check
    n = C.c_int()
    check(lib.dpgo_phase_trace(prob._h, None, 0, C.byref(n)))
    buf = np.zeros((n.value, 16))
    check(lib.dpgo_phase_trace(prob._h, buf.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
    return buf


def main():
    import dpgo_b200
    from bench import lifting_matrix, load_fixture
    name = sys.argv[1] if len(sys.argv) > 1 else "sphere2500"
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    modes = [int(a) for a in sys.argv[3:]] or [2, 0]
    z, d, n = load_fixture(name)
    X0 = np.asfortranarray(lifting_matrix(d, r) @ z["T_chordal"])
    for mode in modes:
        gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r,
                                                 precon_mode=mode)
        prm = dpgo_b200.default_params()
        for _ in range(3):
            Xg, res = gp.optimize(X0, prm)
        busy = trace(gp)
        line = {"dataset": name, "r": r, "precon_mode": gp.precon_mode(), "solve_ms": res["elapsed_ms"],
                "tcg_iters": res["inner_iters"], "barriers": res["n_barriers"], "ctas": int(busy.shape[0]),
                "cost2": 2 * res["f_opt"], "phases": {}}
        for i, ph in enumerate(PHASES):
            tot = res["phase_ms"][i]
            if tot == 0 and not busy[:, i].any():
                continue
            b = busy[:, i]
            line["phases"][ph] = {"cta0_total_ms": round(tot, 4), "busy_max_ms": round(float(b.max()), 4),
                                  "busy_median_ms": round(float(np.median(b)), 4),
                                  "busy_min_ms": round(float(b.min()), 4), "slowest_cta": int(b.argmax())}
        print(json.dumps(line), flush=True)
        gp.close()


if __name__ == "__main__":
    main()
