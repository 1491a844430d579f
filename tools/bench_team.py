"""Multi-agent leg of bench.py: grid3D.g2o split over 8 agents (BASELINE.json configs[2]; torus3D / city10000 with
--team-dataset), agents block-distributed over the ranks (8/N agents per GPU), one process per GPU.

Two parallel block schedules; the reference's own synchronous driver updates ONE agent per
iteration, which cannot use more than one GPU:
  all      every agent optimizes in every round from the poses its neighbours published at the end of the
           previous one -- the equal-rate instance of the reference's asynchronous mode
           (src/PGOAgent.cpp:475-499; no acceleration there, :477).  All N GPUs work in every round: this is
           the series the strong-scaling figure is quoted on (default).
  colored  the agents of one colour of the robot graph optimize, the others do the non-optimizing iterate, with
           Nesterov acceleration (configs[2] to the letter).  A chain robot graph has 2 colours, so at most
           half of the agents work at a time and 8 GPUs cannot be more than 4x one.
A step = one round; the metric counts completed agent updates (PGOAgent::iterate(true)) per second, "scaling":
"strong".  Three series per run:
  device   public poses move inside the C-ABI (dpgo_exchange: device pack, NCCL send/recv over NVLink, on the
           rank's stream), local solves are queued (dpgo_optimize_slot_async): a round is ordered by the stream
           and NCCL only.  L2 is flushed at the start of every round, inside the timed region.  -> `value`
  host     the same rounds through the host-facing calls: poses leave the device into pinned host memory,
           travel between the processes as host buffers (gloo) and are uploaded again; every solve blocks until
           its ROPTResult is on the host (the reference's updateX).  -> `e2e`
  anchor   the whole workload on rank 0's GPU alone (same schedule, device series) in the same run -> the
           denominator of `speedup_vs_1gpu_same_workload`.
  async    (reported beside the others) the reference's asynchronous mode proper: no barrier, no collective, every
           rank iterates its agents at its own rate; public poses are stored into mailboxes in the neighbour's GPU
           memory (CUDA IPC, NVLink peer stores: dpgo_publish) and read as consistent snapshots (dpgo_collect).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def workload(dataset="grid3D", agents=8, r=5, schedule="all"):
    if schedule == "all":
        return (f"{dataset} {agents} agents r={r} all-agents-per-round RBCD (asynchronous-style, no acceleration), "
                "RTR(3 outer, <=50 tCG)")
    return f"{dataset} {agents} agents r={r} sync RBCD + Nesterov, coloured parallel schedule, RTR(3 outer, <=50 tCG)"


def team_config(dataset, agents, r, schedule, gnc=0):
    """The `config` object of the N > 1 lines: identical for our arm and the reference arm."""
    w = workload(dataset, agents, r, schedule)
    if gnc:
        w += f", GNC_TLS loop-closure weights updated every {gnc} rounds"
    return {"workload": w, "dataset": dataset, "agents": agents, "r": r, "schedule": schedule, "gnc_interval": gnc}


def _fixture(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return z, int(z["d"]), int(z["n"])


class Series:
    """One timed series of rounds on a DeviceTeam."""

    def __init__(self, rank, world, local_rank, dataset, agents, r, schedule, mode, stream, flush=None, gnc=0):
        import dpgo_b200
        from dpgo_b200 import rbcd
        from bench import lifting_matrix
        self.rank, self.world, self.mode, self.schedule = rank, world, mode, schedule
        self.z, self.d, self.n = _fixture(dataset)
        z = self.z
        self.r = r
        self.X0 = np.asfortranarray(lifting_matrix(self.d, r) @ z["T_chordal"])
        t0 = time.time()
        self.team = rbcd.DeviceTeam(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], self.n, self.d, r, agents,
                                    device=local_rank, stream=stream, rank=rank, world=world,
                                    acceleration=(schedule != "all"), native_exchange=(mode == "device"),
                                    host_exchange=(mode == "host"), peer_mailboxes=(mode == "async"))
        self.team.set_async(mode in ("device", "async"))
        self.agents_total = agents
        self.setup_s = time.time() - t0
        self.flush = flush
        self.gnc, self.gnc_on, self.weight_updates = int(gnc), False, 0
        self.dpgo_b200 = dpgo_b200

    def reset(self):
        self.team.set_X(self.X0)
        if self.mode == "async":
            self.team.publish_all()          # everybody's initial poses (the callers barrier before the first solve)
        self.team.round = 0
        self.team.host_stats = {"d2h": 0, "h2d": 0}
        for ag in self.team.agents.values():
            ag.iteration = 0
            ag.updates = 0

    def step(self):
        if self.flush is not None:
            self.flush.zero_()
        if self.mode == "async":
            return len(self.team.step_async())       # local agents only: no coordination with the other ranks
        active = self.team.step_all() if self.schedule == "all" else self.team.step_colored()
        if self.mode == "host":       # the step's result on the host: every solve's ROPTResult (blocking updateX)
            for a in active:
                if a in self.team.agents:
                    self.team.agents[a].result()
        if self.gnc_on and self.team.round % self.gnc == 0:
            # GNC robust weights (PGOAgent::updateMeasurementWeights, src/PGOAgent.cpp:1104-1142): residuals of all
            # loop closures on the device, new weights, Q and the exact preconditioner rebuilt on the device
            self.team.update_weights()
            self.weight_updates += 1
        return len(active)

    def launches(self):
        n = sum(ag.prob.launch_count() for ag in self.team.agents.values())
        if self.team.native is not None:
            n += self.team.native.launch_count()
        return n

    def run(self, steps, warmup):
        import torch
        import torch.distributed as dist

        def sync():
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()
                torch.cuda.synchronize()

        self.reset()
        for _ in range(warmup):
            self.step()
        self.reset()
        self.gnc_on = self.gnc > 0       # weight updates belong to the timed rounds only (the warm-up keeps weight 1)
        sync()
        l0 = self.launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        upd = 0
        for _ in range(steps):
            upd += self.step()
        ev1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = ev0.elapsed_time(ev1)
        if self.mode == "host":
            ms = max(ms, wall_ms)        # host work is part of the end-to-end round
        launches = self.launches() - l0
        stats = dict(self.team.host_stats)
        if self.mode == "async":
            upd = self.agents_total * steps          # every agent of every rank did `steps` solves
        if self.world > 1:
            t = torch.tensor([ms, wall_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0]), float(t[1])
            c = torch.tensor([float(launches), float(stats["d2h"]), float(stats["h2d"])], device="cuda",
                             dtype=torch.float64)
            dist.all_reduce(c)
            launches, stats = int(c[0]), {"d2h": int(c[1]), "h2d": int(c[2])}
        X = self.team.assemble()
        return dict(ms=ms, wall_ms=wall_ms, updates=upd, steps=steps, warmup=warmup, value=upd / (ms / 1e3),
                    ms_per_step=ms / steps, launches=launches, X=X, d2h_per_step=stats["d2h"] / steps,
                    h2d_per_step=stats["h2d"] / steps, setup_s=self.setup_s, colors=self.team.colors,
                    owner=self.team.owner, weight_updates=self.weight_updates)

    def close(self):
        self.team.close()


def measure(steps, warmup, rank=0, world=1, local_rank=0, dataset="grid3D", agents=8, r=5, schedule="all",
            mode="device", flush_l2=True, gnc=0):
    """One series (see the module docstring); returns the timed results, identical on every rank."""
    import torch
    torch.cuda.set_device(local_rank)
    tstream = torch.cuda.current_stream()
    if tstream.cuda_stream == 0:         # one explicit stream per rank: library kernels, torch copies, NCCL, events
        tstream = torch.cuda.Stream()
        torch.cuda.set_stream(tstream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if (flush_l2 and mode in ("device", "async")) else None
    s = Series(rank, world, local_rank, dataset, agents, r, schedule, mode, tstream.cuda_stream, flush, gnc)
    try:
        out = s.run(steps, max(warmup, 3))
    finally:
        s.close()
    out["flush_l2"] = flush is not None
    return out


def central_eval(X, dataset, r, local_rank, stream=None):
    """Centralized 2f and Riemannian gradient norm of the assembled iterate (the driver's evaluation,
    examples/MultiRobotExample.cpp:209-226), through the C-ABI."""
    import dpgo_b200
    z, d, n = _fixture(dataset)
    central = dpgo_b200.problem_from_measurements(z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"], n, d, r,
                                                  device=local_rank, stream=stream, build_precon=False)
    try:
        return 2 * central.f(X), central.RieGradNorm(X)
    finally:
        central.close()


def cpu_team_baseline(rounds=2, warmup=2, dataset="grid3D", agents=8, r=5, schedule="all", threads=(1,), gnc=0):
    """The oracle's agents with the same schedule and compiled local solves (oracle/cpu_port), timed once per entry
    of `threads`.  1: one host core, agents one after another (how the reference's MultiRobotExample runs them).
    k > 1: the agents that optimize in the same round run concurrently, one core each (how a one-process-per-robot
    deployment of the reference uses the host); the compiled solve releases the GIL.  The driver's centralized
    evaluation is not part of the timed rounds.  Returns {threads: result}."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pgo, rbcd as orbcd
    z, d, n = _fixture(dataset)
    meas = pgo.make_measurements(d, z["p1"], z["p2"], z["R"], z["t"], z["kappa"], z["tau"])
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    central = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((r, (d + 1) * n)), d)
    state = {}

    def fresh_team():
        team = orbcd.Team(meas, n, agents, r, acceleration=(schedule != "all"))
        for a in team.agents:
            a.use_cpu_port = True
        state["team"], state["colors"] = team, orbcd.robot_graph_coloring(team.agents)

    def one(k, pool, weights=False):
        # same sequence as Team.step_colored / step_all; the optimizing agents (they share no edge / only read
        # the previous round's poses) go to the pool
        team, colors = state["team"], state["colors"]
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
        if weights and gnc > 0 and (k + 1) % gnc == 0:
            team.update_weights()
        return len(active)

    def restart():
        team = state["team"]
        team.set_X(X0)
        for a in team.agents:
            a.iteration = 0

    out = {}
    for th in threads:
        fresh_team()                       # weights back to 1 for every timed series
        restart()
        pool = ThreadPoolExecutor(max_workers=th) if th > 1 else None
        for k in range(max(warmup, 2)):    # also factorizes every agent's preconditioner (both colours)
            one(k, pool)
        restart()
        t0 = time.perf_counter()
        upd = sum(one(k, pool, weights=True) for k in range(rounds))
        dt = time.perf_counter() - t0
        if pool is not None:
            pool.shutdown()
        out[th] = dict(value=upd / dt, ms_per_step=dt / rounds * 1e3, rounds=rounds, warmup=max(warmup, 2),
                       cost2=2 * central.f(state["team"].assemble()), threads=th)
    return out


def parallel_agents(schedule, agents):
    return agents if schedule == "all" else max(1, agents // 2)


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
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    schedule = args.schedule
    ds = dict(dataset=args.team_dataset, agents=args.team_agents, r=args.team_r)
    K, W = args.steps, max(args.warmup, 3)
    gnc = int(getattr(args, "gnc_interval", 0))
    ds_run = dict(ds, gnc=gnc)
    res = measure(K, W, rank, world, local_rank, schedule=schedule, mode="device", **ds_run)
    clocks = sampler.stop() if rank == 0 else None
    warm = measure(K, W, rank, world, local_rank, schedule=schedule, mode="device", flush_l2=False, **ds_run)
    host = measure(K, W, rank, world, local_rank, schedule=schedule, mode="host", **ds_run)
    other = "colored" if schedule == "all" else "all"
    res2 = measure(K, W, rank, world, local_rank, schedule=other, mode="device", **ds_run)
    asyn = None
    if not gnc:       # the asynchronous mode proper (no acceleration, src/PGOAgent.cpp:477)
        try:
            asyn = measure(K, W, rank, world, local_rank, schedule="all", mode="async", **ds)
        except Exception as exc:
            asyn = {"error": repr(exc)}
    # 1-GPU anchors of both schedules on rank 0's GPU (the other ranks wait), same steps
    anchor = anchor2 = None
    if rank == 0:
        anchor = measure(K, W, 0, 1, local_rank, schedule=schedule, mode="device", **ds_run)
        anchor2 = measure(K, W, 0, 1, local_rank, schedule=other, mode="device", **ds_run)
    if world > 1:
        dist.barrier()
    if rank == 0:
        cost2, gradnorm = central_eval(res["X"], ds["dataset"], ds["r"], local_rank, tstream.cuda_stream)
        cost2_other, _ = central_eval(res2["X"], ds["dataset"], ds["r"], local_rank, tstream.cuda_stream)
        same_as_anchor = bool(np.array_equal(res["X"], anchor["X"]))
        host_dev = float(np.max(np.abs(host["X"] - res["X"])))
        par = parallel_agents(schedule, ds["agents"])
        th = max(1, min(par, os.cpu_count() or 1))
        cpus = cpu_team_baseline(min(args.cpu_steps, 2 * gnc if gnc else 4), 2, schedule=schedule,
                                 threads=(1, th) if th > 1 else (1,), gnc=gnc, **ds)
        cpu, cpu_par = cpus[1], cpus[th]
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": f"{ds['dataset']}.g2o (fixture parsed from the reference's data file)",
            "config": team_config(schedule=schedule, gnc=gnc, **ds),
            "details": {
                "gnc": (f"GNC_TLS weight update of all loop closures every {gnc} rounds inside the timed rounds "
                        f"({res['weight_updates']} updates: device residual kernel, host RobustCost, Q and exact "
                        "preconditioner rebuilt on the device)") if gnc else None,
                "n": _fixture(ds["dataset"])[2],
                "agents_per_gpu": ds["agents"] / world, "colors": res["colors"], "owner": res["owner"],
                "step": (f"one round = {ds['agents']} agent updates (iterate(true))" if schedule == "all" else
                         f"one colour round = the agents of one colour ({par} at most) update (iterate(true)), "
                         "the others do the non-optimizing iterate"),
                "l2": "flushed at the start of every round (256 MB device write inside the timed region); the same "
                      "rounds without the flush are under `warm_l2`",
                "exchange": ("dpgo_exchange: device pack + NCCL send/recv of the public poses inside the C-ABI, on the "
                             "rank's stream") if world > 1 else "dpgo_exchange: same-device gathers (single GPU)",
                "host": "stream-ordered rounds: local solves queued with dpgo_optimize_slot_async, no host wait "
                        "inside a round"},
            "warm_l2": {"value": warm["value"], "ms_per_step": warm["ms_per_step"]},
            "speedup_vs_1gpu_same_workload": res["value"] / anchor["value"],
            "anchor_1gpu": {"value": anchor["value"], "ms_per_step": anchor["ms_per_step"], "steps": K,
                            "note": "the same workload, schedule and series with all agents on rank 0's GPU, "
                                    "measured in this run", "poses_identical": same_as_anchor},
            "clocks": clocks,
            "e2e": {"value": host["value"], "unit": UNIT, "ms_per_step": host["ms_per_step"],
                    "h2d_bytes_per_step": host["h2d_per_step"], "d2h_bytes_per_step": host["d2h_per_step"],
                    "note": "the same rounds through the host-facing calls: public poses D2H into pinned host memory, "
                            "moved between the processes as host buffers (gloo), H2D into the receivers; every local "
                            "solve blocks until its ROPTResult is on the host (the reference's updateX)",
                    "max_abs_diff_vs_device_series": host_dev},
            "gpu_launches": int(res["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_rtr_fused (per-agent fused RTR solve)", "achieved": None,
                         "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "see the N=1 line: the dominant kernel and its roofline are measured there"},
            "cpu_baseline": {"value": cpu_par["value"], "unit": UNIT, "cores": cpu_par["threads"], "kind": "port",
                             "ms_per_step": cpu_par["ms_per_step"],
                             "sample": f"{cpu_par['rounds']} rounds ({schedule} schedule) of the oracle's {ds['agents']} agents "
                                       "(compiled C++ local solves, oracle/cpu_port), the agents that optimize in the "
                                       f"same round on one host core each, {os.cpu_count()} host cores visible",
                             "sequential_one_core": {"value": cpu["value"], "ms_per_step": cpu["ms_per_step"],
                                                     "note": "agents one after another on one core, as the "
                                                             "reference's MultiRobotExample runs them"}},
            "parity": {"cost2_after_timed_rounds": cost2, "gradnorm": gradnorm,
                       "poses_identical_to_1gpu_run": same_as_anchor,
                       "cpu_cost2_after_its_rounds": cpu_par["cost2"]},
            "asynchronous_peer_mailboxes": (None if asyn is None else ({"error": asyn["error"]} if "error" in asyn else {
                "value": asyn["value"], "unit": UNIT, "ms_per_step": asyn["ms_per_step"], "steps": K,
                "cost2_after_timed_iterations": central_eval(asyn["X"], ds["dataset"], ds["r"], local_rank,
                                                             tstream.cuda_stream)[0],
                "note": "every agent does `steps` solves at its own rate, no barrier or collective inside the timed "
                        "region; poses through dpgo_publish / dpgo_collect (mailboxes in the neighbour's GPU memory); "
                        "the cost differs from the synchronous series by the staleness of the poses, not by arithmetic"})),
            "other_schedule": {
                "config": team_config(schedule=other, gnc=gnc, **ds), "value": res2["value"], "unit": UNIT,
                "ms_per_step": res2["ms_per_step"], "steps": K, "updates": res2["updates"],
                "cost2_after_timed_rounds": cost2_other,
                "speedup_vs_1gpu_same_workload": res2["value"] / anchor2["value"],
                "anchor_1gpu": {"value": anchor2["value"], "ms_per_step": anchor2["ms_per_step"]},
                "note": "not the headline series: same graph and agents, the other parallel block schedule"},
        }
        (emit or (lambda l: print(json.dumps(l), flush=True)))(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
