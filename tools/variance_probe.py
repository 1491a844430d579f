"""Development probe (GPU): run-to-run spread of one fused solve on sphere2500 under the call
patterns bench.py and tools/dd_probe.py use (own stream vs torch stream, host buffers vs slots,
with / without a large torch allocation made first)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dpgo_b200  # noqa: E402
from bench import lifting_matrix  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "sphere2500.npz"))
d, n, r = int(z["d"]), int(z["n"]), 5
X0 = np.asfortranarray(lifting_matrix(d, r) @ z["T_chordal"])
prm = dpgo_b200.default_params()


def stats(tag, fn, reps=20):
    ms, back = [], []
    for _ in range(3):
        fn()
    for _ in range(reps):
        res = fn()
        ms.append(res["elapsed_ms"])
        back.append(res["phase_ms"][11])
    print(json.dumps({"case": tag, "min": round(min(ms), 3), "mean": round(float(np.mean(ms)), 3),
                      "max": round(max(ms), 3), "back_rhs_min": round(min(back), 3),
                      "back_rhs_mean": round(float(np.mean(back)), 3),
                      "steps": [round(v, 3) for v in ms]}), flush=True)


def make(stream=None):
    return dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r,
                                               device=0, stream=stream)


torch.cuda.set_device(0)
gp = make()
stats("own stream, host buffers", lambda: gp.optimize(X0, prm)[1])
gp.slot_set(dpgo_b200.SLOT_Y, X0)
stats("own stream, slots", lambda: gp.optimize_slot(dpgo_b200.SLOT_Y, prm))
gp.close()
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
gp = make(ts.cuda_stream)
gp.slot_set(dpgo_b200.SLOT_Y, X0)
stats("torch stream, slots", lambda: gp.optimize_slot(dpgo_b200.SLOT_Y, prm))
gp.close()
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
gp = make(ts.cuda_stream)
gp.slot_set(dpgo_b200.SLOT_Y, X0)
stats("torch stream, slots, after a 256 MB torch allocation", lambda: gp.optimize_slot(dpgo_b200.SLOT_Y, prm))


def flushed():
    big.zero_()
    return gp.optimize_slot(dpgo_b200.SLOT_Y, prm)


stats("same, L2 flushed before every solve", flushed)
gp.close()
