"""Run under torchrun with 2 ranks (one GPU each): rounds of the multi-agent schedules with the public poses moved by
dpgo_exchange (NCCL send / recv issued inside the C-ABI) must give the poses of a single rank that holds all agents
(same-device gathers), bit for bit.  Driven by tests/test_gpu_d_async_solve.py::test_native_exchange_across_ranks and
by the GPU sessions of the round."""
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
    name = sys.argv[1] if len(sys.argv) > 1 else "smallGrid3D"
    A = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    meas, n, z = load_dataset(name)
    d, r = meas.d, 5
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    ok = True
    for schedule, rounds in (("colored", 33), ("all", 6)):
        res = {}
        for mode in ("single", "ranks"):
            kw = dict(rank=rank, world=world) if mode == "ranks" else dict(rank=0, world=1)
            team = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A, device=local,
                                   stream=tstream.cuda_stream, acceleration=(schedule == "colored"),
                                   native_exchange=True, **kw)
            team.set_async(True)
            team.set_X(X0)
            for _ in range(rounds):
                (team.step_colored if schedule == "colored" else team.step_all)()
            res[mode] = team.assemble()
            team.close()
        same = bool(np.array_equal(res["single"], res["ranks"]))
        ok = ok and same
        if rank == 0:
            print(f"{name} {A} agents, {schedule}: {world} ranks == 1 rank: {same} "
                  f"(max diff {float(np.max(np.abs(res['single'] - res['ranks']))):.3g})", flush=True)
    dist.barrier()
    if rank == 0 and ok:
        print("native exchange ok", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
