#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's .g2o datasets (run in the build container,
where /root/reference exists; the GPU box only sees the committed fixtures).

Each fixture holds the parsed measurements of one dataset exactly as the reference's parser
produces them (oracle.pgo.read_g2o restates src/DPGO_utils.cpp:113-257): tail/head pose ids,
R, t, kappa, tau, plus the chordal initialization T (d x (d+1)n) computed by the oracle
(oracle.pgo.chordal_initialization restates src/DPGO_solver.cpp:220-269), and a few scalar
known answers of the oracle (cost at the chordal / odometry initial guess) used as golden
values by the tests.

Usage: python tools/make_fixtures.py [names...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import pgo  # noqa: E402

DATA = "/root/reference/data"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
DEFAULT = ["tinyGrid3D", "smallGrid3D", "sphere2500", "torus3D", "grid3D", "city10000"]


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        t0 = time.time()
        meas, n = pgo.read_g2o(os.path.join(DATA, name + ".g2o"))
        d = meas.d
        T = pgo.chordal_initialization(meas, n)
        Q = pgo.connection_laplacian(meas, n)
        prob = pgo.QuadraticProblem(Q, np.zeros((d, (d + 1) * n)), d)
        od = np.where(meas.p2 == meas.p1 + 1)[0]
        Tod = pgo.odometry_initialization(meas.subset(od), n)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            d=np.int64(d), n=np.int64(n),
            p1=meas.p1.astype(np.int32), p2=meas.p2.astype(np.int32),
            R=meas.R, t=meas.t, kappa=meas.kappa, tau=meas.tau,
            T_chordal=T,
            cost2_chordal=np.float64(2 * prob.f(T)),
            cost2_odometry=np.float64(2 * prob.f(Tod)),
            q_scalar_nnz=np.int64(Q.nnz),
        )
        print(f"{name}: d={d} n={n} m={len(meas)} 2f(chordal)={2 * prob.f(T):.10g} "
              f"2f(odom)={2 * prob.f(Tod):.10g} [{time.time() - t0:.1f}s]")


if __name__ == "__main__":
    main(sys.argv[1:] or DEFAULT)
