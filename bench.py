#!/usr/bin/env python
"""bench.py -- RBCD iterations/sec of the B200-native local solve (driver contract in the task
statement).  One "step" = one RBCD iteration = one `QuadraticOptimizer::optimize` call
(ref: src/QuadraticOptimizer.cpp:26-48, called from PGOAgent::updateX src/PGOAgent.cpp:983)
with the reference's default ROptParameters (RTR, 3 outer iterations, <= 50 tCG, radius 100,
gradnorm_tol 1e-2; include/DPGO/DPGO_types.h:53-61) from the lifted chordal initialization.

  N = 1 : BASELINE.json configs[1] -- sphere2500.g2o, one agent, r = 5 (fixture
          tests/golden/sphere2500.npz, parsed from the reference's data file by
          tools/make_fixtures.py).  The line also carries the Q*X roofline at scale (synthetic 262 144- and
          1 000 000-pose grids, `roofline.qx_scale`) and the 1-GPU anchors of the multi-agent workload.
  N > 1 : BASELINE.json configs[2] -- grid3D.g2o split over 8 agents, agents sharded over the N ranks
          (8/N agents per GPU), public poses exchanged with NCCL send/recv inside the C-ABI
          ("scaling": "strong"); see tools/bench_team.py for the two parallel block schedules and the
          device / host / anchor series.  A step = one round; value counts completed agent updates
          (iterate(true)) per second.

  --impl reference : the CPU oracle (the restated reference path, oracle/) timed on the host
          cores of the same box, same config / metric / steps / warmup.
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


SINGLE_WORKLOAD = "sphere2500 1 agent r=5 RTR(3 outer, <=50 tCG) from lifted chordal init"


def single_config():
    """The `config` object of the N = 1 lines: identical for our arm and the reference arm."""
    return {"workload": SINGLE_WORKLOAD, "dataset": "sphere2500", "agents": 1, "r": 5, "schedule": "single agent"}


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
    """The reference's own CPU implementation of the path on this box's host cores: the compiled port of the
    oracle (the reference itself cannot be built here: Eigen / SuiteSparse / ROPTLIB absent), same config, metric,
    steps and warm-up as our arm.  Rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = max(1, args.steps), max(0, args.warmup)
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # same config as our N > 1 arm.  The reference's example driver runs its agents one after another on one
        # thread; a one-process-per-robot deployment runs the agents that optimize in the same round concurrently.
        # The headline value is the latter (all the host threads the path can use), the former is beside it.
        from tools import bench_team
        ds = dict(dataset=args.team_dataset, agents=args.team_agents, r=args.team_r)
        par = bench_team.parallel_agents(args.schedule, args.team_agents)
        th = max(1, min(par, os.cpu_count() or 1))
        cpus = bench_team.cpu_team_baseline(K, W, schedule=args.schedule, threads=(th, 1) if th > 1 else (1,),
                                            gnc=args.gnc_interval, **ds)
        cpu, seq = cpus[th], cpus[1]
        emit({
            "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": cpu["warmup"], "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": f"{args.team_dataset}.g2o (fixture parsed from the reference's data file)",
            "config": bench_team.team_config(schedule=args.schedule, gnc=args.gnc_interval, **ds),
            "details": {"note": CPU_KIND},
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": th, "kind": "port",
                             "sample": f"{K} rounds ({args.schedule} schedule), the {par} agents of a round on one core "
                                       f"each, {os.cpu_count()} host cores visible",
                             "sequential_one_core": {"value": seq["value"], "ms_per_step": seq["ms_per_step"]}},
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "final_cost_2f": cpu["cost2"]})
        return
    val, ms, res = oracle_steps("sphere2500", 5, K, W)
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "sphere2500.g2o (fixture parsed from the reference's data file)",
        "config": single_config(),
        "details": {"note": CPU_KIND + "; the reference itself is single-threaded per agent"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{K} optimize() calls on sphere2500 (r=5), {os.cpu_count()} host cores visible"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_cost_2f": 2 * res.fOpt,
    })


def qx_at_scale(sizes, peak, stream):
    """Q*X achieved HBM GB/s at roofline scale: one entry per grid edge length L (n = L^3 poses)."""
    import dpgo_b200
    from dpgo_b200 import synthetic
    out = []
    for L in sizes:
        g = synthetic.grid3d(L)
        gq = dpgo_b200.problem_from_measurements(g["p1"], g["p2"], g["R"], g["t"], g["kappa"], g["tau"], g["n"], 3, 5,
                                                 device=0, stream=stream, build_precon=False)
        gq.slot_set(0, np.random.default_rng(0).standard_normal((5, 4 * g["n"])))
        nb = gq.bytes_qx()
        us = gq.time_qx(10, True)
        # the product is checked against linearity on the same handle (size-independent property)
        rec = {"n": g["n"], "edges": int(len(g["p1"])), "bytes": nb, "flushed_us": us, "gbs": nb / us / 1e3,
               "frac": nb / us / 1e3 / peak, "bytes_over_time_le_peak": bool(nb / us / 1e3 <= peak)}
        # the per-pose kernels on the same arrays (QF retraction 3, polar projection 4, rounding (r+d)/r tile arrays)
        X = np.asfortranarray(lifting_matrix(3, 5) @ g["T_true"])
        rng = np.random.default_rng(1)
        for slot in range(3):
            gq.slot_set(slot, X if slot == 0 else X + 0.05 * rng.standard_normal(X.shape))
        tile = 5 * 4 * g["n"] * 8.0
        rec["pose_kernels"] = {}
        for op, nm, byts in ((0, "qf_retraction", 3 * tile), (1, "polar_projection", 4 * tile), (2, "rounding", 1.6 * tile)):
            t_us = gq.time_pose_op(op, 10, True)
            rec["pose_kernels"][nm] = {"bytes": byts, "flushed_us": t_us, "frac": byts / t_us / 1e3 / peak}
        out.append(rec)
        gq.close()
    return out


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
                                             n, d, r, device=0, stream=stream,
                                             precon_mode=args.precon_mode)
    prm = dpgo_b200.default_params(fused=1 if args.fused else 0)
    K, W = args.steps, max(args.warmup, 3)
    # L2 is flushed between timed steps (a 256 MB write, outside the per-step event pairs): the
    # two-level preconditioner (58 MB) would otherwise stay L2-resident from one step to the next.
    # Inside a step the solver re-reads it ~36 times; that reuse is part of the step.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed_device_steps(prob, flush_l2=True):
        prob.slot_set(dpgo_b200.SLOT_Y, X0)
        for _ in range(W):
            r_ = prob.optimize_slot(dpgo_b200.SLOT_Y, prm)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        tot = {"n_launches": 0, "n_qx": 0, "n_precon": 0, "n_pose_sweeps": 0, "elapsed_ms": 0.0}
        per_step = []
        for e0, e1 in evs:
            if flush_l2:
                flush.zero_()
            e0.record()
            r_ = prob.optimize_slot(dpgo_b200.SLOT_Y, prm)
            e1.record()
            for k in tot:
                tot[k] += r_[k]
            per_step.append(round(r_["elapsed_ms"], 4))
        torch.cuda.synchronize()
        tot["per_step_ms"] = per_step
        return sum(e0.elapsed_time(e1) for e0, e1 in evs), tot, r_

    # ---- device-resident timing: X0 lives in HBM (slot Y), result stays in HBM (slot X)
    sampler = ClockSampler(0)
    sampler.start()
    ms, tot, res = timed_device_steps(gp)
    clocks = sampler.stop()
    ms_warm, tot_warm, _ = timed_device_steps(gp, flush_l2=False)
    launches, nq, npc, nsw = tot["n_launches"], tot["n_qx"], tot["n_precon"], tot["n_pose_sweeps"]
    value = K / (ms / 1e3)
    mode = gp.precon_mode()

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
    e2e_ms = 0.0
    for _ in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        check(lib.dpgo_optimize(gp._h, C.byref(prm), pin, pout, C.byref(rr)))   # returns after the D2H copy
        e2e_ms += (time.perf_counter() - t0) * 1e3
    e2e_value = K / (e2e_ms / 1e3)
    nbytes = X0.size * 8
    Xg = xout.numpy().T.copy()

    # ---- roofline of the dominant kernel, CUDA events on the same stream
    peak, peak_src = measured_peaks()
    pre_us = gp.time_precon(20, False)
    pre_us_cold = gp.time_precon(10, True)
    pre_bytes = gp.bytes_precon()
    qx_us_warm, qx_us_cold, qx_bytes = gp.time_qx(50, False), gp.time_qx(20, True), gp.bytes_qx()
    # algorithmic bytes of one step = one launch of the fused kernel ("bytes per unit"):
    # every preconditioner application streams its dense blocks once, every Q*X pass reads the
    # block-CSR once (SURVEY 8(d) formula), every per-pose sweep reads 2 and writes 1 lifted array
    sweep_bytes = 3.0 * X0.size * 8

    def step_roofline(ms_, npc_, nq_, nsw_, pre_bytes_):
        sb = (npc_ * pre_bytes_ + nq_ * qx_bytes + nsw_ * sweep_bytes) / K
        return sb, sb / (ms_ / K * 1e-3) / 1e9

    step_bytes, step_gbs = step_roofline(ms, npc, nq, nsw, pre_bytes)
    # dram__bytes_read + dram__bytes_write of this kernel cannot be counted inside a run: the figure is the one of the
    # committed ncu --set full capture of the same command, labelled as such (traffic_source)
    traffic, traffic_src = None, None
    for cand in ("r02_fused_traffic.json", "r01_fused_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", cand)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch_mode%d" % mode, tj.get("dram_bytes_per_launch") if mode == 0 else None)
            traffic_src = f"profiles/{cand} (ncu --set full capture of k_rtr_fused<5,3,{mode}>, not measured in this run)"
            if traffic is not None:
                break
    pre_kernel = {0: "k_precon_gemv<5> (full dense inverse, 800 MB)",
                  2: "k_strip_gemv<5> x3 + k_dd_sep_rhs + k_dd_back_rhs (two-level, 55 MB, L2 resident)"}[mode]
    roofline = {"bound": "hbm",
                "kernel": "k_rtr_fused<5,3,%d> (whole optimize() = 1 launch; dominated by the (Q+0.1I)^-1 apply)" % mode,
                "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                "frac": step_gbs / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": step_bytes, "launch_ms": ms / K,
                "precon_apply_alone": {"kernel": pre_kernel, "bytes": pre_bytes, "us": pre_us,
                                       "us_cold_l2": pre_us_cold,
                                       "achieved": pre_bytes / pre_us / 1e3,
                                       "frac": pre_bytes / pre_us / 1e3 / peak,
                                       "share_of_step": (res["phase_ms"][1] + res["phase_ms"][2]) / max(res["elapsed_ms"], 1e-9),
                                       "share_note": "in-kernel clocks of the last timed step: preconditioner "
                                                     "phases / kernel time"}}
    if mode >= 2:
        roofline["note"] = ("two-level exact preconditioner: 14x fewer algorithmic bytes than the dense inverse and "
                            "L2 resident within a step, so the step is grid-barrier / latency bound, not HBM bound; "
                            "the HBM-bound formulation of the same solve is timed below (dense_inverse_variant)")
        gp0 = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"],
                                                  n, d, r, device=0, stream=stream, precon_mode=0)
        ms0, tot0, res0 = timed_device_steps(gp0)
        b0 = gp0.bytes_precon()
        sb0, gbs0 = step_roofline(ms0, tot0["n_precon"], tot0["n_qx"], tot0["n_pose_sweeps"], b0)
        us0 = gp0.time_precon(20, False)
        roofline["dense_inverse_variant"] = {
            "kernel": "k_rtr_fused<5,3,0>", "ms_per_step": ms0 / K, "value": K / (ms0 / 1e3),
            "algorithmic_bytes_per_launch": sb0, "achieved": gbs0, "frac": gbs0 / peak,
            "precon_apply_alone": {"bytes": b0, "us": us0, "achieved": b0 / us0 / 1e3, "frac": b0 / us0 / 1e3 / peak},
            "final_cost_2f": 2 * res0["f_opt"], "tcg_iters": res0["inner_iters"]}
        gp0.close()
    qx = {"kernel": "k_qx<5,3> on sphere2500", "bytes": qx_bytes, "warm_us": qx_us_warm, "warm_gbs": qx_bytes / qx_us_warm / 1e3,
          "cold_l2_us": qx_us_cold, "cold_l2_gbs": qx_bytes / qx_us_cold / 1e3,
          "cold_l2_frac_of_peak": qx_bytes / qx_us_cold / 1e3 / peak,
          "note": "sphere2500 Q*X moves 2.4 MB: L2-resident and launch-bound in situ; see roofline-scale run (tools/qx_scale.py)"}

    roofline["qx_in_situ"] = qx
    if args.qx_scale:
        # the block-CSR Q*X on inputs that stream from HBM (synthetic 3-D grids, see
        # dpgo_b200/synthetic.py), L2 flushed before every timed launch, CUDA events on the launching stream
        gp.close()
        gp = None
        roofline["qx_scale"] = qx_at_scale([int(v) for v in args.qx_scale.split(",")], peak, stream)

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu_val, cpu_ms, cres = oracle_steps(name, r, args.cpu_steps, 1)
    gap = abs(res["f_opt"] - cres.fOpt) / abs(cres.fOpt)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "sphere2500.g2o (fixture parsed from the reference's data file)",
        "config": single_config(),
        "details": {"n": n, "d": d, "r": r, "solver": "fused persistent kernel" if args.fused else "one launch per op",
                    "l2": "flushed between timed steps (256 MB device write outside the per-step CUDA-event pairs)",
                    "preconditioner": {0: "full dense inverse",
                                       2: "two-level (nested dissection + Schur complement)"}[mode],
                    "outer_iters": res["outer_iters"], "tcg_iters": res["inner_iters"],
                    "qx_per_step": nq / K, "precon_per_step": npc / K},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
        "gpu_launches": int(launches),
        # the same K steps again without the L2 flush, and the part of a step spent inside the solver
        # kernel (CUDA events around the launch inside the library; the rest is the call's launch +
        # result read-back)
        "ms_per_step_warm_l2": ms_warm / K,
        "solver_kernel_ms_per_step": {"l2_flushed": tot["elapsed_ms"] / K, "l2_warm": tot_warm["elapsed_ms"] / K,
                                      "l2_flushed_steps": tot["per_step_ms"]},
        "fused_phase_ms": dict(zip(["cost_grad", "precon_stream", "precon_finish", "hessvec", "tcg_update",
                                    "tcg_direction", "retract_copy", "unused", "dd_interior_y", "dd_sep_rhs",
                                    "dd_schur", "dd_back_rhs", "dd_interior_w"], res.get("phase_ms", []))),
        "grid_barriers_per_step": res.get("n_barriers", 0),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                         "ms_per_step": cpu_ms,
                         "sample": f"{args.cpu_steps} optimize() calls, same workload, compiled C++ port (oracle/cpu_port), "
                                   f"{os.cpu_count()} host cores visible (reference is single-threaded per agent)"},
        "parity": {"final_cost_2f_gpu": 2 * res["f_opt"], "final_cost_2f_cpu": 2 * cres.fOpt,
                   "rel_gap": gap, "e2e_result_matches": bool(abs(2 * rr.f_opt - 2 * res["f_opt"]) < 1e-9),
                   "x_finite": bool(np.isfinite(Xg).all())},
    }
    # chordal initialization of the same dataset (ref: src/DPGO_solver.cpp:220-269): the device
    # path through the C-ABI with host buffers, beside the oracle's sparse direct solves on the host
    try:
        import dpgo_b200
        zc = load_fixture(name)[0]
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            Tc, cinfo = dpgo_b200.chordal_initialization(zc["p1"], zc["p2"], zc["R"], zc["t"], zc["kappa"], zc["tau"], n, d)
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        line["chordal_initialization"] = {
            "dataset": name, "device_ms": best, "cg_iterations": [cinfo["rotation_iterations"], cinfo["translation_iterations"]],
            "relative_residuals": [cinfo["rotation_residual"], cinfo["translation_residual"]],
            "note": "device_ms is the whole call: two handles (Q assembly + preconditioner set-up) and two CG solves; "
                    "the CPU side (oracle's sparse direct solves) is cpu_baseline.chordal_initialization"}
        # CPU leg: the oracle's chordal initialization (sparse LU on the host) on the same measurements, and the check
        from oracle import pgo as _pgo
        t0 = time.perf_counter()
        To = _pgo.chordal_initialization(_pgo.make_measurements(d, zc["p1"], zc["p2"], zc["R"], zc["t"], zc["kappa"], zc["tau"]), n)
        line["cpu_baseline"]["chordal_initialization"] = {
            "oracle_sparse_lu_ms": (time.perf_counter() - t0) * 1e3,
            "device_rel_diff_vs_oracle": float(np.linalg.norm(Tc - To) / np.linalg.norm(To))}
    except Exception as exc:
        line["chordal_initialization"] = {"error": repr(exc)}
    if args.example:
        # the reference's own acceptance driver (examples/MultiRobotExample.cpp, unmodified) through the C++ drop-in
        try:
            from tools import bench_example
            line["multi_robot_example"] = bench_example.measure()
        except Exception as exc:
            line["multi_robot_example"] = {"error": repr(exc)}
    if args.team_steps > 0:
        # 1-GPU anchors of the strong-scaling series (BASELINE configs[2]): the rows the N > 1 lines divide by
        if gp is not None:
            gp.close()
        from tools import bench_team
        ds = dict(dataset=args.team_dataset, agents=args.team_agents, r=args.team_r)
        for sched in ("all", "colored"):
            try:
                t = bench_team.measure(args.team_steps, 3, 0, 1, 0, schedule=sched, mode="device", **ds)
                c2, gn = bench_team.central_eval(t["X"], ds["dataset"], ds["r"], 0, stream)
                line[f"{ds['dataset']}_{ds['agents']}agents_{sched}"] = {
                    "config": bench_team.team_config(schedule=sched, **ds), "n_gpus": 1, "value": t["value"],
                    "unit": UNIT, "ms_per_step": t["ms_per_step"], "steps": t["steps"],
                    "cost2_after_timed_rounds": c2, "gradnorm": gn,
                    "series": "device (dpgo_exchange, stream-ordered rounds, L2 flushed every round)"}
            except Exception as exc:       # the headline line above must survive a failure here
                line[f"{ds['dataset']}_{ds['agents']}agents_{sched}"] = {"error": repr(exc)}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--schedule", default="all", choices=["colored", "all"],
                    help="multi-agent series (N > 1): every agent in every round without acceleration (the "
                         "asynchronous-style parallel schedule, all GPUs busy: default) or the coloured parallel "
                         "schedule with Nesterov acceleration (BASELINE configs[2] to the letter; both are measured, "
                         "this picks the headline)")
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--precon-mode", type=int, default=None,
                    help="storage of the exact preconditioner: 0 dense inverse, 2 two-level (default: the library's "
                         "choice by size)")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--team-dataset", default="grid3D", help="multi-agent series (N > 1): fixture in tests/golden")
    ap.add_argument("--team-agents", type=int, default=8)
    ap.add_argument("--team-r", type=int, default=5)
    ap.add_argument("--team-steps", type=int, default=10,
                    help="rounds of the multi-agent 1-GPU anchors appended to the N=1 line (0 = skip)")
    ap.add_argument("--example", type=int, default=1,
                    help="N = 1: also time the reference's MultiRobotExample driver through the C++ drop-in (0 = skip)")
    ap.add_argument("--gnc-interval", type=int, default=0,
                    help="multi-agent series: GNC_TLS robust weight update of all loop closures every this many "
                         "rounds inside the timed rounds (BASELINE configs[4]: --team-dataset city10000 --team-agents 4 "
                         "--team-r 3 --gnc-interval 5); 0 = plain least squares")
    ap.add_argument("--qx-scale", default="64,100",
                    help="edge lengths L of the synthetic L^3-pose grids for the Q*X roofline at scale ('' = skip)")
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
