"""Timing of the reference's own acceptance driver through the C++ drop-in: examples/MultiRobotExample.cpp compiled
UNMODIFIED against dpgo_b200/host/include (dpgo_b200/host/bin/multi-robot-example, built where the reference tree is
present), run on the CUDA path.  The driver prints one line per iteration (greedy block selection, r = 5, Nesterov
acceleration, centralized evaluation every iteration; examples/MultiRobotExample.cpp:170-247); the arrival time of
every line gives iterations / s without touching the driver.  CPU arm beside it: the oracle's restatement of the same
driver (oracle/rbcd.py Team.step_greedy with the compiled local solves), same schedule, same evaluation."""
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LINE = re.compile(r"Iter = (\d+) \| robot = (\d+) \| cost = ([-+.\deE]+) \| gradnorm = ([-+.\deE]+)")


def _write_fixture(dataset, path):
    from util_g2o import write_g2o
    z = np.load(os.path.join(ROOT, "tests", "golden", dataset + ".npz"))
    write_g2o(path, int(z["d"]), z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"])


def run_example(dataset, robots, workdir, max_seconds=120.0):
    """Runs the driver binary; returns iterations/s over the printed iterations (first line excluded: it carries
    the set-up of the preconditioners), the last cost and gradient norm, and the whole wall time."""
    from dpgo_b200 import build
    exe = build.host_binary("multi-robot-example")
    path = os.path.join(workdir, dataset + ".g2o")
    _write_fixture(dataset, path)
    t_start = time.perf_counter()
    proc = subprocess.Popen([exe, str(robots), path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    stamps, rows = [], []
    try:
        for ln in proc.stdout:
            m = LINE.search(ln)
            if m:
                stamps.append(time.perf_counter())
                rows.append((int(m.group(1)), int(m.group(2)), float(m.group(3)), float(m.group(4))))
            if time.perf_counter() - t_start > max_seconds:
                proc.kill()
                break
    finally:
        proc.wait()
    wall = time.perf_counter() - t_start
    if len(rows) < 3:
        return {"error": f"driver printed {len(rows)} iterations (rc {proc.returncode})"}
    its = rows[-1][0] - rows[0][0]
    dt = stamps[-1] - stamps[0]
    return {"dataset": dataset, "robots": robots, "iterations": rows[-1][0] + 1, "iterations_per_s": its / dt,
            "ms_per_iteration": dt / its * 1e3, "first_iteration_at_s": stamps[0] - t_start, "wall_s": wall,
            "final_cost_2f": rows[-1][2], "final_gradnorm": rows[-1][3], "stopped": "gradnorm < 0.1" if rows[-1][3] < 0.1
            else ("1000 iterations" if rows[-1][0] >= 999 else "time limit"),
            "agent_sequence_head": [r[1] for r in rows[:12]]}


def run_cpu_driver(dataset, robots, iterations, r=5):
    """The oracle's restatement of the same driver on one host core (compiled local solves)."""
    from oracle import pgo, rbcd as orbcd
    z = np.load(os.path.join(ROOT, "tests", "golden", dataset + ".npz"))
    d, n = int(z["d"]), int(z["n"])
    meas = pgo.make_measurements(d, z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"])
    team = orbcd.Team(meas, n, robots, r, acceleration=True)
    for a in team.agents:
        a.use_cpu_port = True
    team.set_X(pgo.lifting_matrix(d, r) @ z["T_chordal"])
    team.step_greedy()                       # first iteration: factorizations
    t0 = time.perf_counter()
    s = None
    for _ in range(iterations):
        s = team.step_greedy()
    dt = time.perf_counter() - t0
    return {"iterations": iterations, "iterations_per_s": iterations / dt, "ms_per_iteration": dt / iterations * 1e3,
            "cost_2f_after": s["cost"], "gradnorm_after": s["gradnorm"], "cores": 1, "kind": "port"}


def measure(cases=(("smallGrid3D", 5, 30), ("grid3D", 8, 12))):
    import tempfile
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for dataset, robots, cpu_iters in cases:
            rec = run_example(dataset, robots, tmp)
            if "error" not in rec:
                rec["cpu_same_driver"] = run_cpu_driver(dataset, robots, cpu_iters)
                rec["speedup_vs_cpu_driver"] = rec["iterations_per_s"] / rec["cpu_same_driver"]["iterations_per_s"]
            out.append(rec)
    return out


if __name__ == "__main__":
    import json
    for rec in measure():
        print(json.dumps(rec), flush=True)
