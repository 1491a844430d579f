"""Run under torchrun with N ranks (one GPU each): the asynchronous publication through peer memory across processes
(CUDA IPC mailboxes, NVLink peer stores).  (1) in lockstep it must reproduce the all-agents schedule of a single rank
bit for bit; (2) free-running (no barrier inside the timed loop) it must converge: centralized cost after K solves per
agent within 1e-3 of the lockstep cost, finite, and decreasing from the initial cost.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_dataset
    from dpgo_b200 import rbcd
    from oracle import pgo
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    name = sys.argv[1] if len(sys.argv) > 1 else "torus3D"
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    meas, n, z = load_dataset(name)
    d, r = meas.d, 5
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    central = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((r, (d + 1) * n)), d)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    def make(**kw):
        t = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A, device=local,
                            stream=tstream.cuda_stream, acceleration=False, **kw)
        t.set_async(True)
        t.set_X(X0)
        return t

    single = make(rank=0, world=1, native_exchange=True)
    for _ in range(K):
        single.step_all()
    X_all = single.assemble()
    single.close()
    lock = make(rank=rank, world=world, peer_mailboxes=True)
    lock.publish_all()
    barrier()
    for _ in range(K):
        lock.step_async_lockstep(barrier)
    X_lock = lock.assemble()
    lock.close()
    same = bool(np.array_equal(X_lock, X_all))
    free = make(rank=rank, world=world, peer_mailboxes=True)
    free.publish_all()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(K):
        free.step_async()
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    barrier()
    X_free = free.assemble()
    free.close()
    c0, c_all, c_free = 2 * central.f(X0), 2 * central.f(X_all), 2 * central.f(X_free)
    ok = same and np.isfinite(X_free).all() and c_free < c0 and abs(c_free - c_all) <= 1e-3 * c_all
    if rank == 0:
        print(json.dumps({"dataset": name, "agents": A, "ranks": world, "solves_per_agent": K,
                          "lockstep_equals_all_schedule": same, "cost2_initial": c0, "cost2_all_schedule": c_all,
                          "cost2_free_running": c_free, "free_running_ms": float(ms[0]),
                          "free_running_updates_per_s": A * K / (float(ms[0]) / 1e3), "ok": bool(ok)}), flush=True)
        if ok:
            print("peer mailboxes ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
