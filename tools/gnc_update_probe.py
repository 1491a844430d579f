#!/usr/bin/env python
"""Where a GNC weight update of one agent spends its time (city10000 / 4 agents, BASELINE configs[4]):
residual kernel + read-back, host RobustCost, Q re-weighting, preconditioner set-up.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import lifting_matrix, load_fixture
    from dpgo_b200 import rbcd
    from dpgo_b200.robust import RobustCost
    name, A, r = (sys.argv[1] if len(sys.argv) > 1 else "city10000"), 4, 3
    z, d, n = load_fixture(name)
    team = rbcd.DeviceTeam(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r, A, acceleration=False,
                           native_exchange=True)
    team.set_async(True)
    team.set_X(lifting_matrix(d, r) @ z["T_chordal"])
    for _ in range(5):
        team.step_all()
    team.exchange(list(range(A)))
    ag = team.agents[0]
    rc = RobustCost()

    def timed(fn, reps=5):
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best * 1e3, out

    t_err, (ep, es) = timed(lambda: ag.prob.measurement_errors(rbcd.SLOT_X, ag.prob.neighbor_buffer(0)))
    t_w, (wp, ws) = timed(lambda: (np.where(ag.spec.priv_fixed, 1.0, rc.weights(np.sqrt(ep))),
                                   np.where(ag.spec.shared_fixed, 1.0, rc.weights(np.sqrt(es)))))
    t_q, _ = timed(lambda: ag.prob.update_weights(wp, ws, False))
    t_qp, _ = timed(lambda: ag.prob.update_weights(wp, ws, True))
    print(json.dumps({"dataset": name, "agent_poses": ag.spec.n, "private_edges": int(len(wp)), "shared_edges": int(len(ws)),
                      "residuals_ms": t_err, "robust_cost_host_ms": t_w, "update_weights_Q_only_ms": t_q,
                      "update_weights_Q_and_precon_ms": t_qp, "precon_ms": t_qp - t_q}), flush=True)
    team.close()


if __name__ == "__main__":
    main()
