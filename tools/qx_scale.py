#!/usr/bin/env python
"""Roofline-scale measurement of the block-CSR Q*X kernel: synthetic 3-D
grid pose graphs large enough that Q and X do not fit in the 126 MB L2, timed with CUDA events
on the launching stream.  Prints one JSON line per size and writes gpurun_out/qx_scale.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpgo_b200  # noqa: E402
from dpgo_b200 import synthetic  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [40, 64, 100]
    peak = 6481.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    out = []
    for L in sizes:
        t0 = time.time()
        g = synthetic.grid3d(L)
        tg = time.time() - t0
        t0 = time.time()
        gp = dpgo_b200.problem_from_measurements(g["p1"], g["p2"], g["R"], g["t"], g["kappa"], g["tau"],
                                                 g["n"], 3, 5, build_precon=False)
        tb = time.time() - t0
        rng = np.random.default_rng(0)
        gp.slot_set(0, rng.standard_normal((5, 4 * g["n"])))
        nbytes = gp.bytes_qx()
        Xh = rng.standard_normal((5, 4 * g["n"]))
        rec = dict(L=L, n=g["n"], m=len(g["p1"]), bytes=nbytes, gen_s=tg, build_s=tb, peak_gbs=peak, variants=[])
        ref = None
        # 0 = lane-group kernel, 1 = lane-group kernel + L2 prefetch (distance 4096), 2 = shared-memory staged
        # X tiles (phase_qx_tiles), 3 = lane-group kernel, two blocks per step with the column indices one step ahead
        for variant, dist in ((0, 0), (1, 4096), (2, 0), (3, 0), (-1, 0)):
            gp.set_qx_variant(variant, dist)
            prod = gp.qx(Xh)
            if ref is None:
                ref = prod
            err = float(np.linalg.norm(prod - ref) / np.linalg.norm(ref))
            w1, c1 = gp.time_qx(20, False), gp.time_qx(10, True)
            rec["variants"].append(dict(variant=variant, distance=dist, rel_diff_vs_variant0=err,
                                        back_to_back_us=w1, back_to_back_gbs=nbytes / w1 / 1e3, flushed_us=c1,
                                        flushed_gbs=nbytes / c1 / 1e3, frac_of_measured_peak=nbytes / c1 / 1e3 / peak))
        gp.set_qx_variant(-1, 0)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        gp.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "qx_scale.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
