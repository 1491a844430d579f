#!/bin/bash
# Round-2 GPU session 31 (1 GPU): compute-sanitizer (memcheck, then racecheck) over a small run of every device path:
# smoke() (fused solver, dense form), a two-level solve, chordal initialization, a GNC weight refresh, the staged
# per-pose kernels.
O=gpurun_out/s31
mkdir -p $O
cat > $O/san_run.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as g
import dpgo_b200
from bench import lifting_matrix
g.smoke()
z = np.load("tests/golden/smallGrid3D.npz"); d, n, r = int(z["d"]), int(z["n"]), 5
gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r, precon_mode=2)
X0 = np.asfortranarray(lifting_matrix(d, r) @ z["T_chordal"])
X, res = gp.optimize(X0, dpgo_b200.default_params()); print("two-level solve", res["outer_iters"], res["inner_iters"], 2 * res["f_opt"])
m = len(z["p1"]); w = np.random.default_rng(0).uniform(0.2, 1.0, m)
gp.update_weights(w, None, True); X, res = gp.optimize(X0, dpgo_b200.default_params()); print("after weight refresh", 2 * res["f_opt"])
print("retract", np.linalg.norm(gp.retract(X, 0.01 * X)), "polar", np.linalg.norm(gp.project_manifold(X + 0.01)), "round", gp.round_trajectory().shape)
gp.close()
T, info = dpgo_b200.chordal_initialization(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d); print("chordal", info)
PY
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python $O/san_run.py > $O/$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|two-level|chordal|after weight" $O/$tool.log | tail -8
done
