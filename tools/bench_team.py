"""Multi-agent leg of bench.py: BASELINE.json configs[2] -- grid3D.g2o split over 8 agents,
synchronous RBCD with Nesterov acceleration, r = 5, coloured parallel block schedule, agents
block-distributed over the ranks (8/N agents per GPU), public poses over NCCL send/recv.
A step = one colour round (4 agents optimize, 4 do the non-optimizing iterate); the metric counts
completed agent updates (PGOAgent::iterate(true)) per second -> "scaling": "strong"."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKLOAD = "grid3D 8 agents r=5 sync RBCD + Nesterov, coloured parallel schedule, RTR(3 outer, <=50 tCG)"
# --schedule all: every agent optimizes every round from the previous round's poses, no acceleration
# (the equal-rate instance of the reference's asynchronous mode, BASELINE configs[3]); all GPUs busy
WORKLOAD_ALL = "grid3D 8 agents r=5 all-agents-per-round RBCD (asynchronous-style, no acceleration), RTR(3 outer, <=50 tCG)"


def workload(dataset="grid3D", agents=8, r=5, schedule="colored"):
    if schedule == "all":
        return (f"{dataset} {agents} agents r={r} all-agents-per-round RBCD (asynchronous-style, no acceleration), "
                "RTR(3 outer, <=50 tCG)")
    return f"{dataset} {agents} agents r={r} sync RBCD + Nesterov, coloured parallel schedule, RTR(3 outer, <=50 tCG)"


assert workload() == WORKLOAD and workload(schedule="all") == WORKLOAD_ALL


def _fixture(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return z, int(z["d"]), int(z["n"])


def measure(steps, warmup, rank=0, world=1, local_rank=0, dataset="grid3D", agents=8, r=5, e2e=True,
            schedule="colored", async_rounds=True):
    """Returns a dict with the timed results (identical on every rank)."""
    import torch
    import torch.distributed as dist
    import dpgo_b200
    from dpgo_b200 import rbcd
    from bench import lifting_matrix
    torch.cuda.set_device(local_rank)
    # one explicit stream per rank: library kernels, torch device copies, NCCL stream
    # dependencies and the timing events are all ordered on it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    z, d, n = _fixture(dataset)
    X0 = np.asfortranarray(lifting_matrix(d, r) @ z["T_chordal"])
    t0 = time.time()
    team = rbcd.DeviceTeam(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r, agents,
                           device=local_rank, stream=stream, rank=rank, world=world,
                           acceleration=(schedule != "all"))
    setup_s = time.time() - t0
    step = team.step_all if schedule == "all" else team.step_colored

    def reset():
        team.set_X(X0)
        team.round = 0
        for ag in team.agents.values():
            ag.iteration = 0
            ag.updates = 0

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(nsteps, with_host_eval, central=None):
        sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        upd = 0
        cost = None
        for _ in range(nsteps):
            active = step()
            upd += len(active)
            if with_host_eval:   # the driver's per-iteration evaluation: getX of every agent + f
                X = team.assemble()
                if central is not None:
                    cost = 2 * central.f(X)
        ev1.record()
        torch.cuda.synchronize()
        ms = max(ev0.elapsed_time(ev1), 0.0)
        wall_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms, wall_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0]), float(t[1])
        return ms, wall_ms, upd, cost

    def preflight_async():
        """One stream-ordered solve per local agent, no exchange involved; every rank learns whether
        it worked everywhere (a rank that failed alone would leave the others waiting in NCCL)."""
        ok, why = 1, None
        try:
            for ag in team.agents.values():
                ag.prob.optimize_slot_async(rbcd.SLOT_X, ag.params)
                ag.prob.optimize_result()
        except Exception as exc:
            ok, why = 0, repr(exc)
        if world > 1:
            t = torch.tensor([ok], device="cuda", dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = int(t[0])
        return bool(ok), why

    def series(async_mode):
        team.set_async(async_mode)
        reset()
        for _ in range(W):
            step()
        reset()
        l0 = sum(ag.prob.launch_count() for ag in team.agents.values())
        ms_, wall_, upd_, _ = timed(steps, False)
        # launches of OUR kernels in the timed region (counted by the library per handle, this rank)
        launches_ = sum(ag.prob.launch_count() for ag in team.agents.values()) - l0
        return ms_, wall_, upd_, launches_, team.assemble()

    W = max(warmup, 3)
    # Two series over the same rounds.  "blocking": PGOAgent::updateX as in the reference, the host
    # waits for every local solve.  "stream-ordered": the solve is queued (dpgo_optimize_slot_async),
    # the host runs ahead of the device and only the stream / NCCL order the round.  Same kernels, same
    # inputs: the poses after the timed rounds must be identical, otherwise the blocking series is the
    # one reported.
    reset()
    ms, wall_ms, upd, launches, X = series(False)
    blocking = dict(value=upd / (ms / 1e3), ms_per_step=ms / steps)
    mode, mode_note, stream_ordered = "blocking", None, None
    if async_rounds:
        ok, why = preflight_async()
        if ok:
            try:
                ms_a, wall_a, upd_a, launches_a, Xa = series(True)
                dev = float(np.max(np.abs(Xa - X))) if Xa.shape == X.shape else float("inf")
            except Exception as exc:     # keep the blocking series (a sticky CUDA error would end the run anyway)
                ms_a, wall_a, upd_a, launches_a, Xa, dev = ms, wall_ms, -1, launches, X, float("inf")
                why = repr(exc)
            if upd_a == upd and dev <= 1e-12 * max(1.0, float(np.max(np.abs(X)))):
                stream_ordered = dict(value=upd_a / (ms_a / 1e3), ms_per_step=ms_a / steps)
                if ms_a <= ms:
                    ms, wall_ms, launches, X = ms_a, wall_a, launches_a, Xa
                    mode, mode_note = "stream-ordered", f"max |X - X_blocking| = {dev:.3g}"
                else:
                    team.set_async(False)
                    mode_note = (f"stream-ordered series identical (max |X - X_blocking| = {dev:.3g}) but not faster: "
                                 f"{stream_ordered['ms_per_step']:.4f} ms/step")
            else:
                team.set_async(False)
                mode_note = f"stream-ordered series rejected: max |X - X_blocking| = {dev:.3g}" + (f" ({why})" if why else "")
        else:
            mode_note = f"stream-ordered solve unavailable: {why}"
    central = None
    cost = gradnorm = None
    if rank == 0:
        central = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"],
                                                      n, d, r, device=local_rank, stream=stream,
                                                      build_precon=False)
        cost = 2 * central.f(X)
        gradnorm = central.RieGradNorm(X)
    out = dict(ms=ms, wall_ms=wall_ms, updates=upd, steps=steps, warmup=W, setup_s=setup_s,
               value=upd / (ms / 1e3), ms_per_step=ms / steps, cost2=cost, gradnorm=gradnorm,
               n=n, d=d, r=r, agents=agents, colors=team.colors, owner=team.owner,
               fused_launches=launches, host_mode=mode, host_mode_note=mode_note, blocking=blocking,
               stream_ordered=stream_ordered)
    if e2e:
        reset()
        e_ms, e_wall, e_upd, e_cost = timed(steps, True, central)
        e_ms = max(e_ms, e_wall)
        out.update(e2e_value=e_upd / (e_ms / 1e3), e2e_ms_per_step=e_ms / steps,
                   e2e_bytes=int(X.size * 8))
    if central is not None:
        central.close()
    team.close()
    return out


def cpu_team_baseline(rounds=2, dataset="grid3D", agents=8, r=5, schedule="colored", threads=(1,)):
    """The oracle's agents with the same schedule and compiled local solves (oracle/cpu_port), timed
    once per entry of `threads`.  1: one host core, agents one after another (how the reference's
    MultiRobotExample runs them).  k > 1: the agents that optimize in the same round run
    concurrently, one core each (how a one-process-per-robot deployment of the reference would
    use the host); the compiled solve releases the GIL.  The driver's centralized evaluation is
    not part of the timed rounds.  Returns {threads: result}."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pgo, rbcd as orbcd
    z, d, n = _fixture(dataset)
    meas = pgo.make_measurements(d, z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"])
    team = orbcd.Team(meas, n, agents, r, acceleration=(schedule != "all"))
    for a in team.agents:
        a.use_cpu_port = True
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    colors = orbcd.robot_graph_coloring(team.agents)

    def one(k, pool):
        # same sequence as Team.step_colored / step_all; the optimizing agents (they share no
        # edge / only read the previous round's poses) go to the pool
        active = list(range(agents)) if schedule == "all" else colors[k % len(colors)]
        for a in team.agents:
            if a.id not in active:
                a.iterate(False)
        for a in team.agents:
            if a.id in active:
                team._exchange_to(a)
        todo = [a for a in team.agents if a.id in active]
        if pool is None:
            for a in todo:
                a.iterate(True)
        else:
            list(pool.map(lambda a: a.iterate(True), todo))
        return len(active)

    def restart():
        team.set_X(X0)
        for a in team.agents:
            a.iteration = 0

    restart()
    one(0, None)                      # warm-up rounds: factorize every agent's preconditioner
    one(1, None)
    out = {}
    for th in threads:
        restart()
        pool = ThreadPoolExecutor(max_workers=th) if th > 1 else None
        t0 = time.perf_counter()
        upd = sum(one(k, pool) for k in range(rounds))
        dt = time.perf_counter() - t0
        if pool is not None:
            pool.shutdown()
        out[th] = dict(value=upd / dt, ms_per_step=dt / rounds * 1e3, rounds=rounds,
                       cost2=2 * team.central.f(team.assemble()), threads=th)
    return out


def run(args, emit=None):
    import torch
    import torch.distributed as dist
    from bench import METRIC, UNIT, ClockSampler
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at level VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    schedule = getattr(args, "schedule", "colored")
    # BASELINE configs[2] unless asked otherwise (configs[3]: --team-dataset torus3D --schedule all;
    # configs[4] without the weight updates: --team-dataset city10000 --team-agents 4 --team-r 3)
    ds = dict(dataset=getattr(args, "team_dataset", "grid3D"), agents=getattr(args, "team_agents", 8),
              r=getattr(args, "team_r", 5))
    res = measure(args.steps, args.warmup, rank, world, local_rank, schedule=schedule, **ds)
    clocks = sampler.stop() if rank == 0 else None
    # second series for the record: the other parallel schedule of SURVEY 8(e) on the same graph
    other = "colored" if schedule == "all" else "all"
    try:
        res2 = measure(args.steps, args.warmup, rank, world, local_rank, schedule=other, e2e=False, **ds)
    except Exception as exc:  # the headline series above must survive a failure here
        res2 = {"error": repr(exc)}
    if rank == 0:
        # agents that optimize in the same round
        par = res["agents"] if schedule == "all" else max(len(c) for c in res["colors"])
        cpus = cpu_team_baseline(args.cpu_steps if args.cpu_steps < 4 else 2, schedule=schedule,
                                 threads=(1, min(par, os.cpu_count() or 1)), **ds)
        cpu, cpu_par = cpus[1], cpus[max(cpus)]
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": res["steps"],
            "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": f"{ds['dataset']}.g2o (fixture parsed from the reference's data file)",
            "config": {"workload": workload(schedule=schedule, **ds), "n": res["n"], "agents": res["agents"],
                       "agents_per_gpu": res["agents"] / world, "colors": res["colors"], "owner": res["owner"],
                       "step": (f"one round = {res['agents']} agent updates (iterate(true))" if schedule == "all" else
                                f"one colour round = the agents of one colour ({par} at most) update (iterate(true)), "
                                "the others do the non-optimizing iterate"),
                       "l2": "per-agent preconditioner streamed per apply (grid3D/8: two-level, 21.5 MB); the agents "
                             "of a GPU alternate",
                       "exchange": "NCCL send/recv of packed public poses (X and aux Y)" if world > 1 else
                                   "device-to-device copies (single GPU)",
                       "host": res["host_mode"] + " rounds" + (
                           " (local solves queued with dpgo_optimize_slot_async: no host wait inside a round)"
                           if res["host_mode"] == "stream-ordered" else
                           " (the host waits for every local solve, as the reference's updateX does)"),
                       "host_note": res["host_mode_note"]},
            "blocking_updateX": dict(res["blocking"], unit=UNIT,
                                     note="same rounds with the host waiting for every local solve"),
            "stream_ordered_rounds": res["stream_ordered"],
            "clocks": clocks,
            "e2e": {"value": res.get("e2e_value"), "unit": UNIT, "ms_per_step": res.get("e2e_ms_per_step"),
                    "h2d_bytes_per_step": res.get("e2e_bytes"), "d2h_bytes_per_step": res.get("e2e_bytes"),
                    "note": "driver-style round: every agent's X copied to the host and the centralized "
                            "cost evaluated through the C-ABI with host buffers each round"},
            "gpu_launches": int(res["fused_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_rtr_fused (per-agent fused RTR solve)", "achieved": None,
                         "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "see the N=1 line: the dominant kernel and its roofline are measured there"},
            "cpu_baseline": {"value": cpu["value"], "unit": UNIT, "cores": 1, "kind": "port",
                             "ms_per_step": cpu["ms_per_step"],
                             "sample": f"{cpu['rounds']} rounds ({schedule} schedule) of the oracle's 8 agents (compiled C++ local solves, oracle/cpu_port), "
                                       "sequential on one core as the reference's MultiRobotExample runs them",
                             "agents_in_parallel": {"value": cpu_par["value"], "cores": cpu_par["threads"],
                                                    "ms_per_step": cpu_par["ms_per_step"],
                                                    "note": "the agents that optimize in the same round on one host core each"}},
            "parity": {"cost2_after_timed_rounds": res["cost2"], "gradnorm": res["gradnorm"]},
            "other_schedule": ({"error": res2["error"]} if "error" in res2 else {
                "workload": workload(schedule=other, **ds), "value": res2["value"], "unit": UNIT,
                "ms_per_step": res2["ms_per_step"], "steps": res2["steps"], "updates": res2["updates"],
                "cost2_after_timed_rounds": res2["cost2"], "gradnorm": res2["gradnorm"],
                "host": res2["host_mode"], "blocking_updateX": res2["blocking"],
                "note": "not the headline series: same graph and agents, the other parallel block schedule"}),
        }
        (emit or (lambda l: print(json.dumps(l), flush=True)))(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
