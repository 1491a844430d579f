"""CPU test of the N > 1 host logic (world_size 2, gloo): agent -> rank ownership, the per-pair
pack lists and the send/recv pairing of dpgo_b200.rbcd.exchange_poses deliver to every active
agent exactly the neighbour poses the reference's PoseDict exchange would
(getSharedPoseDict / updateNeighborPoses, src/PGOAgent.cpp:97-146, 650-702)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _CpuAgent:
    """Stand-in with the same exchange interface as rbcd.DeviceAgent, on CPU tensors."""

    def __init__(self, spec, X, Y, tile):
        self.spec, self.tile = spec, tile
        self.X, self.Y = X, Y                      # (n, tile) rows = poses
        nslots = max(len(spec.nbr_keys), 1)
        self.nbr = torch.zeros(nslots * tile, dtype=torch.float64)
        self.nbr_aux = torch.zeros(nslots * tile, dtype=torch.float64)
        self.send_idx, self.send_buf, self.send_buf_aux = {}, {}, {}

    def prepare_send(self, b, frames):
        self.send_idx[b] = torch.as_tensor(np.asarray(frames, dtype=np.int64))
        self.send_buf[b] = torch.empty(len(frames) * self.tile, dtype=torch.float64)
        self.send_buf_aux[b] = torch.empty(len(frames) * self.tile, dtype=torch.float64)

    def pack_for(self, b):
        self.send_buf[b].copy_(self.X[self.send_idx[b]].reshape(-1))
        self.send_buf_aux[b].copy_(self.Y[self.send_idx[b]].reshape(-1))

    def recv_view(self, b, aux):
        lo, hi = self.spec.nbr_range[b]
        return (self.nbr_aux if aux else self.nbr)[lo * self.tile:hi * self.tile]

    def host_buffer(self, key, numel):
        if not hasattr(self, "_host"):
            self._host = {}
        if key not in self._host:
            self._host[key] = torch.empty(numel, dtype=torch.float64)
        return self._host[key]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_dataset
    from dpgo_b200 import rbcd
    meas, n, z = load_dataset("smallGrid3D")
    d, r, A = 3, 5, 5
    tile = r * (d + 1)
    ranges, specs = rbcd.build_specs(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, A)
    owner = rbcd.block_owner(A, world)
    rng = np.random.default_rng(123)                 # same on both ranks
    Xs = {a: torch.from_numpy(rng.standard_normal((specs[a].n, tile))) for a in range(A)}
    Ys = {a: torch.from_numpy(rng.standard_normal((specs[a].n, tile))) for a in range(A)}
    agents = {a: _CpuAgent(specs[a], Xs[a], Ys[a], tile) for a in range(A) if owner[a] == rank}
    for a, ag in agents.items():
        for b in specs[a].neighbors:
            ag.prepare_send(b, specs[b].nbr_frames[a])
    colors = rbcd.color_robot_graph(specs)
    ok = True
    for active in colors + [[2]]:
        rbcd.exchange_poses(agents, specs, owner, rank, active, True)
        for a in active:
            if owner[a] != rank:
                continue
            for slot, (b, f) in enumerate(specs[a].nbr_keys):
                got = agents[a].nbr[slot * tile:(slot + 1) * tile]
                got_aux = agents[a].nbr_aux[slot * tile:(slot + 1) * tile]
                ok &= bool(torch.equal(got, Xs[b][f])) and bool(torch.equal(got_aux, Ys[b][f]))
    # every agent active at once (DeviceTeam.step_all / update_weights): both directions of every
    # robot-graph edge travel in the same batch, X only (no acceleration in that schedule)
    for ag in agents.values():
        ag.nbr.zero_(); ag.nbr_aux.zero_()
    everyone = list(range(A))
    rbcd.exchange_poses(agents, specs, owner, rank, everyone, False)
    for a in everyone:
        if owner[a] != rank:
            continue
        for slot, (b, f) in enumerate(specs[a].nbr_keys):
            ok &= bool(torch.equal(agents[a].nbr[slot * tile:(slot + 1) * tile], Xs[b][f]))
        ok &= bool(torch.count_nonzero(agents[a].nbr_aux) == 0)
    # the host-staged form of the same exchange (end-to-end series of bench.py): identical deliveries, and the
    # bytes it reports are the tiles this rank sent (D2H) and received (H2D)
    stats = {}
    for active, acc in [(c, True) for c in colors] + [(everyone, False)]:
        for ag in agents.values():
            ag.nbr.zero_(); ag.nbr_aux.zero_()
        rbcd.exchange_poses_host(agents, specs, owner, rank, active, acc, None, stats)
        for a in active:
            if owner[a] != rank:
                continue
            for slot, (b, f) in enumerate(specs[a].nbr_keys):
                ok &= bool(torch.equal(agents[a].nbr[slot * tile:(slot + 1) * tile], Xs[b][f]))
                if acc:
                    ok &= bool(torch.equal(agents[a].nbr_aux[slot * tile:(slot + 1) * tile], Ys[b][f]))
    want_h2d = sum((2 if acc else 1) * len(specs[a].nbr_keys) * tile * 8
                   for active, acc in [(c, True) for c in colors] + [(everyone, False)] for a in active if owner[a] == rank)
    ok &= stats.get("h2d", 0) == want_h2d and stats.get("d2h", 0) > 0
    ret[rank] = (ok, owner, colors)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        ok, owner, colors = ret[rank]
        assert ok
        assert owner == [0, 0, 0, 1, 1]
        assert colors == [[0, 2, 4], [1, 3]]


def test_partition_matches_reference_driver():
    """examples/MultiRobotExample.cpp:71-119 on a toy chain: contiguous equal split, last robot
    takes the remainder; odometry / private / shared classification."""
    sys.path.insert(0, ROOT)
    from dpgo_b200 import rbcd
    n, A = 11, 3
    p1 = np.array(list(range(n - 1)) + [0, 2, 4]); p2 = np.array(list(range(1, n)) + [9, 3, 10])
    m = len(p1)
    R = np.tile(np.eye(3), (m, 1, 1)); t = np.zeros((m, 3))
    ranges, specs = rbcd.build_specs(p1, p2, R, t, np.ones(m), np.ones(m), n, 3, A)
    assert ranges == [(0, 3), (3, 6), (6, 11)]
    assert [s.n for s in specs] == [3, 3, 5]
    assert specs[0].neighbors == [1, 2] and specs[2].neighbors == [0, 1]
    # edge 2->3 is shared between robots 0 and 1 (local frames 2 and 0)
    assert (1, 0) in specs[0].nbr_keys and (0, 2) in specs[1].nbr_keys
    assert len(specs[0].priv["p1"]) == 2 and len(specs[2].priv["p1"]) == 4
