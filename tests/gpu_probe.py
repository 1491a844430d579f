"""First-contact script for a GPU box: timings of the individual kernels and of optimize()."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from oracle import pgo  # noqa: E402

out = {}
for name, r in (("smallGrid3D", 5), ("sphere2500", 5), ("torus3D", 5), ("grid3D", 5)):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    d, n = int(z["d"]), int(z["n"])
    t0 = time.time()
    gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"],
                                             n, d, r, build_precon=(n <= 8000))
    setup = time.time() - t0
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    gp.slot_set(0, X0)
    rec = dict(n=n, setup_s=setup, qx_us_warm=gp.time_qx(50, False), qx_us_cold=gp.time_qx(20, True),
               qx_bytes=gp.bytes_qx())
    rec["qx_gbs_warm"] = rec["qx_bytes"] / rec["qx_us_warm"] / 1e3
    rec["qx_gbs_cold"] = rec["qx_bytes"] / rec["qx_us_cold"] / 1e3
    if n <= 8000:
        rec["precon_us"] = gp.time_precon(10, False)
        rec["precon_us_cold"] = gp.time_precon(5, True)
        rec["precon_bytes"] = gp.bytes_precon()
        rec["precon_gbs"] = rec["precon_bytes"] / rec["precon_us"] / 1e3
        for fused in (0, 1):
            for rep in range(3):
                t0 = time.time()
                X, res = gp.optimize(X0, dpgo_b200.default_params(fused=fused))
                wall = time.time() - t0
            rec[f"opt_fused{fused}"] = dict(res, wall_ms=wall * 1e3)
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
    gp.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
