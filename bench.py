#!/usr/bin/env python
"""bench.py -- RBCD iterations/sec of the B200-native local solve (driver contract in the task
statement).  One "step" = one RBCD iteration = one `QuadraticOptimizer::optimize` call
(ref: src/QuadraticOptimizer.cpp:26-48, called from PGOAgent::updateX src/PGOAgent.cpp:983)
with the reference's default ROptParameters (RTR, 3 outer iterations, <= 50 tCG, radius 100,
gradnorm_tol 1e-2; include/DPGO/DPGO_types.h:53-61) from the lifted chordal initialization.

  N = 1 : BASELINE.json configs[1] -- sphere2500.g2o, one agent, r = 5 (fixture
          tests/golden/sphere2500.npz, parsed from the reference's data file by
          tools/make_fixtures.py).
  N > 1 : BASELINE.json configs[2] -- grid3D.g2o split over 8 agents, synchronous RBCD with
          Nesterov acceleration, coloured parallel block schedule, agents sharded over the N
          ranks (8/N agents per GPU), public poses exchanged with NCCL send/recv
          ("scaling": "strong").  A step = one colour round; value counts completed agent
          updates (iterate(true)) per second.

  --impl reference : the CPU oracle (the restated reference path, oracle/) timed on the host
          cores of the same box, same config / metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

METRIC = "rbcd_iterations_per_sec"
UNIT = "iterations/s"

_JSON_FD = None


def capture_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries (NCCL's version banner, cuSOLVER,
    verbose solvers) also write to fd 1, so fd 1 is pointed at stderr for the duration of the
    run and the JSON line goes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, int(z["d"]), int(z["n"])


def lifting_matrix(d, r):
    """Deterministic element of St(d, r) (stand-in for fixedStiefelVariable,
    ref: src/DPGO_utils.cpp:488-493; any fixed Stiefel element gives the same objective)."""
    A = np.array([[np.cos(1.0 + 0.7 * i + 1.3 * j) + (1.0 if i == j else 0.0) for j in range(d)]
                  for i in range(r)])
    Q, R = np.linalg.qr(A)
    s = np.sign(np.diag(R)); s[s == 0] = 1
    return Q * s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle on the host cores
# ---------------------------------------------------------------------------------------------
CPU_KIND = ("compiled C++ port of the reference path (oracle/cpu_port: block-CSR Q*X, exact block sparse "
            "Cholesky preconditioner with minimum-degree ordering, RTR/tCG), -O3, single thread like the "
            "reference's per-agent solve")


class _Res:
    pass


def oracle_steps(name, r, steps, warmup):
    """Time optimize() of the CPU oracle (compiled port) on the host: the reference's CPU path."""
    from oracle import pgo
    from oracle.cpu_port import CpuProblem
    z, d, n = load_fixture(name)
    meas = pgo.make_measurements(d, z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"])
    prob = CpuProblem(pgo.connection_laplacian(meas, n), np.zeros((r, (d + 1) * n)), d)
    prob.factorize()                             # set-up (the reference factorizes once per Q too)
    X0 = lifting_matrix(d, r) @ z["T_chordal"]
    res = None
    for _ in range(warmup):
        _, res = prob.optimize(X0)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, res = prob.optimize(X0)
    dt = time.perf_counter() - t0
    out = _Res()
    out.fOpt, out.outer, out.inner = res["f_opt"], res["outer"], res["inner"]
    return steps / dt, dt / steps * 1e3, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # same config as our N > 1 arm: grid3D, 8 agents, coloured schedule, agents run one after
        # another on one host thread (as the reference's MultiRobotExample runs them)
        from tools import bench_team
        rounds = max(1, min(args.steps, 4))
        cpu = bench_team.cpu_team_baseline(rounds)
        emit({
            "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": rounds, "warmup": 2, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "grid3D.g2o (fixture parsed from the reference's data file)",
            "config": {"workload": bench_team.WORKLOAD, "note": CPU_KIND},
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{rounds} colour rounds (4 agent updates each), {os.cpu_count()} host cores visible"},
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "final_cost_2f": cpu["cost2"]})
        return
    steps = max(1, min(args.steps, 100))
    warm = max(1, min(args.warmup, 2))
    name, r = ("sphere2500", 5)
    val, ms, res = oracle_steps(name, r, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "sphere2500.g2o (fixture parsed from the reference's data file)",
        "config": {"workload": "sphere2500 1 agent r=5 RTR(3 outer, <=50 tCG) from lifted chordal init",
                   "note": CPU_KIND + "; "
                           "the reference itself is single-threaded per agent"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{steps} optimize() calls on sphere2500 (r=5), {os.cpu_count()} host cores visible"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_cost_2f": 2 * res.fOpt,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# N = 1: sphere2500, single agent
# ---------------------------------------------------------------------------------------------
def bench_single(args):
    import torch
    import dpgo_b200
    torch.cuda.set_device(0)
    # one explicit stream for everything: the library's kernels are launched on it and
    # torch.cuda.Event (which only sees torch's current stream) times the same stream
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    name, r = "sphere2500", 5
    z, d, n = load_fixture(name)
    X0 = np.asfortranarray(lifting_matrix(d, r) @ z["T_chordal"])
    gp = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"],
                                             n, d, r, device=0, stream=stream)
    prm = dpgo_b200.default_params(fused=1 if args.fused else 0)
    gp.slot_set(dpgo_b200.SLOT_Y, X0)
    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident timing: X0 lives in HBM (slot Y), result stays in HBM (slot X)
    for _ in range(W):
        res = gp.optimize_slot(dpgo_b200.SLOT_Y, prm)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    nq = npc = nsw = 0
    ev0.record()
    for _ in range(K):
        res = gp.optimize_slot(dpgo_b200.SLOT_Y, prm)
        launches += res["n_launches"]; nq += res["n_qx"]; npc += res["n_precon"]; nsw += res["n_pose_sweeps"]
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    value = K / (ms / 1e3)

    # ---- end to end through the public call with HOST buffers (pinned): H2D of X0 and D2H of X
    xin = torch.from_numpy(np.ascontiguousarray(X0.T)).pin_memory()     # same bytes as col-major r x N
    xout = torch.empty_like(xin).pin_memory()
    import ctypes as C
    from dpgo_b200._lib import lib, RoptResult, check
    dp = C.POINTER(C.c_double)
    pin, pout = C.cast(xin.data_ptr(), dp), C.cast(xout.data_ptr(), dp)
    rr = RoptResult()
    for _ in range(W):
        check(lib.dpgo_optimize(gp._h, C.byref(prm), pin, pout, C.byref(rr)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(K):
        check(lib.dpgo_optimize(gp._h, C.byref(prm), pin, pout, C.byref(rr)))
    ev1.record()
    torch.cuda.synchronize()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    e2e_value = K / (e2e_ms / 1e3)
    nbytes = X0.size * 8
    Xg = xout.numpy().T.copy()

    # ---- roofline of the dominant kernel (dense preconditioner apply), CUDA events on the same stream
    peak, peak_src = measured_peaks()
    pre_us = gp.time_precon(20, False)
    pre_bytes = gp.bytes_precon()
    qx_us_warm, qx_us_cold, qx_bytes = gp.time_qx(50, False), gp.time_qx(20, True), gp.bytes_qx()
    # algorithmic bytes of one step = one launch of the fused kernel ("bytes per unit"):
    # every preconditioner application streams the dense inverse once (N^2*8 + 2 r N 8), every
    # Q*X pass reads the block-CSR once (SURVEY 8(d) formula), every per-pose sweep reads 2 and
    # writes 1 lifted array
    sweep_bytes = 3.0 * X0.size * 8
    step_bytes = (npc * pre_bytes + nq * qx_bytes + nsw * sweep_bytes) / K
    step_s = ms / K * 1e-3
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_fused_traffic.json")
    if os.path.exists(tpath):     # dram__bytes_read+write of k_rtr_fused from the committed ncu --set full capture
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm",
                "kernel": "k_rtr_fused<5,3> (whole optimize() = 1 launch; dominated by the dense (Q+0.1I)^-1 apply)",
                "achieved": step_bytes / step_s / 1e9, "peak": peak, "unit": "GB/s",
                "frac": step_bytes / step_s / 1e9 / peak, "peak_source": peak_src, "traffic": traffic,
                "algorithmic_bytes_per_launch": step_bytes, "launch_ms": ms / K,
                "precon_apply_alone": {"kernel": "k_precon_gemv<5>", "bytes": pre_bytes, "us": pre_us,
                                       "achieved": pre_bytes / pre_us / 1e3,
                                       "frac": pre_bytes / pre_us / 1e3 / peak,
                                       "share_of_step": (npc / K) * pre_us / (ms / K * 1e3)}}
    qx = {"bytes": qx_bytes, "warm_us": qx_us_warm, "warm_gbs": qx_bytes / qx_us_warm / 1e3,
          "cold_l2_us": qx_us_cold, "cold_l2_gbs": qx_bytes / qx_us_cold / 1e3,
          "cold_l2_frac_of_peak": qx_bytes / qx_us_cold / 1e3 / peak,
          "note": "sphere2500 Q*X moves 2.4 MB: L2-resident and launch-bound in situ; see roofline-scale run (tools/qx_scale.py)"}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu_val, cpu_ms, cres = oracle_steps(name, r, args.cpu_steps, 1)
    gap = abs(res["f_opt"] - cres.fOpt) / abs(cres.fOpt)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "sphere2500.g2o (fixture parsed from the reference's data file)",
        "config": {"workload": "sphere2500 1 agent r=5 RTR(3 outer, <=50 tCG) from lifted chordal init",
                   "n": n, "d": d, "r": r, "solver": "fused persistent kernel" if args.fused else "one launch per op",
                   "l2": "inputs larger than L2 (dense preconditioner 800 MB streamed every apply)",
                   "outer_iters": res["outer_iters"], "tcg_iters": res["inner_iters"],
                   "qx_per_step": nq / K, "precon_per_step": npc / K},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
        "gpu_launches": int(launches),
        "fused_phase_ms": dict(zip(["cost_grad", "precon_gemv", "precon_finish", "hessvec", "tcg_update",
                                    "tcg_direction", "retract_copy", "unused"], res.get("phase_ms", []))),
        "grid_barriers_per_step": res.get("n_barriers", 0),
        "roofline": roofline, "qx": qx,
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                         "ms_per_step": cpu_ms,
                         "sample": f"{args.cpu_steps} optimize() calls, same workload, compiled C++ port (oracle/cpu_port), "
                                   f"{os.cpu_count()} host cores visible (reference is single-threaded per agent)"},
        "parity": {"final_cost_2f_gpu": 2 * res["f_opt"], "final_cost_2f_cpu": 2 * cres.fOpt,
                   "rel_gap": gap, "e2e_result_matches": bool(abs(2 * rr.f_opt - 2 * res["f_opt"]) < 1e-9),
                   "x_finite": bool(np.isfinite(Xg).all())},
    }
    if args.team_steps > 0:
        # strong-scaling series (BASELINE configs[2]) measured at this N as well, so that the
        # 1 -> 2 -> 4 -> 8 GPU rows of the multi-agent workload have a 1-GPU anchor
        gp.close()
        from tools import bench_team
        t = bench_team.measure(args.team_steps, 3, 0, 1, 0)
        line["grid3D_8agents"] = {
            "workload": bench_team.WORKLOAD, "n_gpus": 1, "value": t["value"], "unit": UNIT,
            "ms_per_step": t["ms_per_step"], "steps": t["steps"], "e2e_value": t.get("e2e_value"),
            "cost2_after_timed_rounds": t["cost2"], "gradnorm": t["gradnorm"]}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--team-steps", type=int, default=10,
                    help="colour rounds of the grid3D/8-agent series appended to the N=1 line (0 = skip)")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- dpgo_b200 has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return bench_single(args)
    from tools import bench_team
    return bench_team.run(args, emit)


if __name__ == "__main__":
    main()
