#!/bin/bash
# Round-2 GPU session 32 (1 GPU): ncu --set full of the dense-inverse form of the fused solver (the HBM-bound
# formulation of the same solve); compute-sanitizer initcheck over the run of session 31.
O=gpurun_out/s32
mkdir -p $O
cat > $O/dense_run.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import dpgo_b200
from bench import lifting_matrix
z = np.load("tests/golden/sphere2500.npz"); d, n, r = int(z["d"]), int(z["n"]), 5
gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r, precon_mode=0)
X0 = np.asfortranarray(lifting_matrix(d, r) @ z["T_chordal"])
for _ in range(3):
    X, res = gp.optimize(X0, dpgo_b200.default_params())
print("dense form", res["elapsed_ms"], 2 * res["f_opt"], res["inner_iters"])
PY
timeout 400 ncu --set full --clock-control none -k regex:k_rtr_fused -s 2 -c 1 -o $O/dense_full -f python $O/dense_run.py > $O/ncu_dense.log 2>&1
ncu -i $O/dense_full.ncu-rep --page raw --csv > $O/dense_full_raw.csv 2>/dev/null
rm -f $O/dense_full.ncu-rep
python tools/ncu_digest.py full $O/dense_full_raw.csv $O/dense_fused_traffic.json k_rtr_fused | tail -1
cp gpurun_out/s31/san_run.py $O/san_run.py 2>/dev/null || sed -n '/^cat > \$O\/san_run.py/,/^PY$/p' tools/sessions/r02_s31.sh | sed '1d;$d' > $O/san_run.py
timeout 300 compute-sanitizer --tool initcheck --print-limit 20 python $O/san_run.py > $O/initcheck.log 2>&1; echo "initcheck rc=$?"
grep -E "ERROR SUMMARY|smoke ok|chordal" $O/initcheck.log | tail -4
