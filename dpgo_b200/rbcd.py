"""Multi-agent synchronous RBCD on device-resident agents (one dpgo_handle per agent).

Host-side mirror of the reference's driver and agent loop for the hot path:
  * partition / edge classification: examples/MultiRobotExample.cpp:71-119
  * PGOAgent::iterate, updateX, Nesterov updates, periodic restart:
    src/PGOAgent.cpp:376-432, 880-995
  * public-pose exchange (getSharedPoseDict / updateNeighborPoses and the aux variants):
    src/PGOAgent.cpp:97-146, 650-702 -- here packed device buffers moved with NCCL send/recv
    (torch.distributed) between ranks, or device-to-device copies inside one rank.
All per-pose arithmetic runs behind the C-ABI (CUDA); this module only sequences calls.

Block schedule: the reference's synchronous driver updates ONE agent per iteration (greedy).
To use more than one GPU the agents of one colour of the robot graph are updated concurrently
(they share no edge, so this equals updating them one after another).
"""
import math

import numpy as np

from .api import SLOT_V, SLOT_X, SLOT_XPREV, SLOT_Y, DeviceProblem, default_params


def partition(p1, p2, n, num_robots):
    """Contiguous equal split, last robot takes the remainder
    (examples/MultiRobotExample.cpp:71-88).  Returns (ranges, robot_of_pose, local_idx)."""
    per = n // num_robots
    if per <= 0:
        raise ValueError("more robots than poses")
    starts = [k * per for k in range(num_robots)]
    ends = [(k + 1) * per for k in range(num_robots)]
    ends[-1] = n
    rob = np.minimum(np.arange(n) // per, num_robots - 1)
    loc = np.arange(n) - np.asarray(starts)[rob]
    return list(zip(starts, ends)), rob, loc


class AgentSpec:
    """Everything agent `a` needs, computed once on every rank (cheap, deterministic)."""

    def __init__(self, a, d, ranges, rob, loc, p1, p2, R, t, kappa, tau):
        self.id = a
        self.n = ranges[a][1] - ranges[a][0]
        r1, r2 = rob[p1], rob[p2]
        l1, l2 = loc[p1], loc[p2]
        priv = np.where((r1 == a) & (r2 == a))[0]
        sh = np.where(((r1 == a) | (r2 == a)) & (r1 != r2))[0]
        self.priv = dict(p1=l1[priv], p2=l2[priv], R=R[priv], t=t[priv], kappa=kappa[priv], tau=tau[priv])
        # the reader's fixedWeight flag (consecutive GLOBAL pose ids, src/DPGO_utils.cpp:178,232)
        self.priv_fixed = (p1[priv] + 1 == p2[priv])
        self.shared_fixed = (p1[sh] + 1 == p2[sh])
        out = r1[sh] == a                               # my pose is the tail (m.r1 == id_)
        my = np.where(out, l1[sh], l2[sh])
        nb_r = np.where(out, r2[sh], r1[sh])
        nb_f = np.where(out, l2[sh], l1[sh])
        # neighbour slots: PoseGraph::nbr_shared_pose_ids_ is a std::set ordered by (robot, frame)
        keys = sorted(set(zip(nb_r.tolist(), nb_f.tolist())))
        slot_of = {k: i for i, k in enumerate(keys)}
        self.nbr_keys = keys
        self.shared = dict(my_idx=my, nbr_slot=np.array([slot_of[(int(r_), int(f_))] for r_, f_ in zip(nb_r, nb_f)],
                                                        dtype=np.int32),
                           outgoing=out.astype(np.uint8), R=R[sh], t=t[sh], kappa=kappa[sh], tau=tau[sh])
        self.neighbors = sorted(set(k[0] for k in keys))
        # contiguous slot range [lo, hi) of each neighbour robot and the frames it must send
        self.nbr_range, self.nbr_frames = {}, {}
        for b in self.neighbors:
            idx = [i for i, k in enumerate(keys) if k[0] == b]
            self.nbr_range[b] = (idx[0], idx[-1] + 1)
            self.nbr_frames[b] = np.array([keys[i][1] for i in idx], dtype=np.int32)
        self.public = sorted(set(my.tolist()))


def color_robot_graph(specs):
    color = {}
    for s in specs:
        used = {color[b] for b in s.neighbors if b in color}
        c = 0
        while c in used:
            c += 1
        color[s.id] = c
    ncol = max(color.values()) + 1 if color else 1
    return [[s.id for s in specs if color[s.id] == c] for c in range(ncol)]


class DeviceAgent:
    """Device-resident PGOAgent state: X, Y, V, XPrev slots + Nesterov scalars."""

    def __init__(self, spec, d, r, num_robots, device, stream, acceleration=True, restart_interval=30,
                 params=None):
        import torch
        self.spec, self.d, self.r, self.id = spec, d, r, spec.id
        self.num_robots = num_robots
        self.acceleration = acceleration
        self.restart_interval = restart_interval
        self.params = params if params is not None else default_params()
        self.prob = DeviceProblem(spec.n, d, r, device, stream)
        p, s = spec.priv, spec.shared
        self.prob.set_private_edges(p["p1"], p["p2"], p["R"], p["t"], p["kappa"], p["tau"])
        self.prob.set_shared_edges(s["my_idx"], s["nbr_slot"], s["outgoing"], s["R"], s["t"], s["kappa"],
                                   s["tau"], num_nbr_slots=len(spec.nbr_keys))
        self.prob.finalize(True)
        tile = r * (d + 1)
        dev = torch.device("cuda", device)
        nslots = max(len(spec.nbr_keys), 1)
        self.nbr = torch.zeros(nslots * tile, dtype=torch.float64, device=dev)       # neighbours' X
        self.nbr_aux = torch.zeros(nslots * tile, dtype=torch.float64, device=dev)   # neighbours' Y
        self.tile = tile
        # native exchange (dpgo_exchange): neighbour poses live in the handle's own buffers
        self.native = False
        self.iteration = 0
        self.gamma = self.alpha = 0.0
        self.last_result = None
        # True: updateX queues the fused solve on the stream and returns (dpgo_optimize_slot_async);
        # the result block is fetched on demand by result().  The reference's updateX blocks.
        self.async_solve = False
        self.updates = 0
        self.w_private = np.ones(len(p["p1"]))
        self.w_shared = np.ones(len(s["my_idx"]))
        # send side: for each neighbour b, which of my frames it needs (device index lists)
        self.send_idx, self.send_buf, self.send_buf_aux = {}, {}, {}

    def prepare_send(self, b, frames):
        import torch
        dev = self.nbr.device
        self.send_idx[b] = torch.as_tensor(np.asarray(frames, dtype=np.int32), device=dev)
        self.send_buf[b] = torch.empty(len(frames) * self.tile, dtype=torch.float64, device=dev)
        self.send_buf_aux[b] = torch.empty(len(frames) * self.tile, dtype=torch.float64, device=dev)

    def set_X(self, X):                                   # PGOAgent::setX + initializeAcceleration
        self.prob.slot_set(SLOT_X, X)
        if self.acceleration:
            for s in (SLOT_XPREV, SLOT_V, SLOT_Y):
                self.prob.slot_copy(s, SLOT_X)
            self.gamma = self.alpha = 0.0

    def get_X(self):
        return self.prob.slot_get(SLOT_X)

    def pack_for(self, b):
        """Fill the send buffers for neighbour b with my X (and Y) poses it needs."""
        idx = self.send_idx[b]
        self.prob.gather_tiles_dev(SLOT_X, idx.numel(), idx.data_ptr(), self.send_buf[b].data_ptr())
        if self.acceleration:
            self.prob.gather_tiles_dev(SLOT_Y, idx.numel(), idx.data_ptr(), self.send_buf_aux[b].data_ptr())

    def host_buffer(self, key, numel):
        """Pinned host staging buffer of the host-staged exchange (allocated at first use)."""
        import torch
        if not hasattr(self, "_host"):
            self._host = {}
        if key not in self._host:
            self._host[key] = torch.empty(numel, dtype=torch.float64, pin_memory=torch.cuda.is_available())
        return self._host[key]

    def recv_view(self, b, aux):
        lo, hi = self.spec.nbr_range[b]
        buf = self.nbr_aux if aux else self.nbr
        return buf[lo * self.tile:hi * self.tile]

    def _update_X(self, do_opt, acceleration):            # PGOAgent::updateX :938-995
        if not do_opt:
            if acceleration:
                self.prob.slot_copy(SLOT_X, SLOT_Y)
            return
        if self.native:
            self.prob.use_neighbor_poses(1 if acceleration else 0)   # setNeighborPoses + constructG on device
        else:
            buf = self.nbr_aux if acceleration else self.nbr
            self.prob.set_neighbor_poses_dev(buf.data_ptr())
        src = SLOT_Y if acceleration else SLOT_X
        if self.async_solve:
            self.prob.optimize_slot_async(src, self.params)
            self.last_result = None
        else:
            self.last_result = self.prob.optimize_slot(src, self.params)
        self.updates += 1

    def result(self):
        """ROPTResult of the last local solve (waits for the stream in asynchronous mode)."""
        if self.last_result is None and self.async_solve and self.updates > 0:
            self.last_result = self.prob.optimize_result()
        return self.last_result

    def update_measurement_weights(self, robust):
        """PGOAgent::updateMeasurementWeights (src/PGOAgent.cpp:1104-1142, robustOptNumResets = 0):
        residual of every loop closure at X and the neighbours' X by one device pass
        (dpgo_measurement_errors), new weights for the edges whose weight is not fixed, GNC
        schedule step, Q and preconditioner rebuilt on the device, acceleration restarted."""
        ep, es = self.prob.measurement_errors(SLOT_X, self.prob.neighbor_buffer(0) if self.native
                                              else self.nbr.data_ptr())
        wp = np.where(self.spec.priv_fixed, self.w_private, robust.weights(np.sqrt(ep)))
        ws = np.where(self.spec.shared_fixed, self.w_shared, robust.weights(np.sqrt(es)))
        self.w_private, self.w_shared = wp, ws
        robust.update()
        self.prob.update_weights(wp, ws, True)
        if self.acceleration:                             # initializeAcceleration :899-908
            for s in (SLOT_XPREV, SLOT_V, SLOT_Y):
                self.prob.slot_copy(s, SLOT_X)
            self.gamma = self.alpha = 0.0

    def iterate(self, do_opt):                            # PGOAgent::iterate :376-432
        self.iteration += 1
        self.prob.slot_copy(SLOT_XPREV, SLOT_X)
        if not self.acceleration:
            self._update_X(do_opt, False)
            return
        R = self.num_robots
        self.gamma = (1 + math.sqrt(1 + 4 * R * R * self.gamma * self.gamma)) / (2 * R)   # :910-914
        self.alpha = 1 / (self.gamma * R)                                                # :916-920
        self.prob.nesterov_update_Y(self.alpha)
        self._update_X(do_opt, True)
        self.prob.nesterov_update_V(self.gamma)
        if (self.iteration + 1) % self.restart_interval == 0:                            # :880-897
            self.prob.slot_copy(SLOT_X, SLOT_XPREV)
            self._update_X(do_opt, False)
            self.prob.slot_copy(SLOT_V, SLOT_X)
            self.prob.slot_copy(SLOT_Y, SLOT_X)
            self.gamma = self.alpha = 0.0


def exchange_poses(agents, specs, owner, rank, active, acceleration):
    """Move the public poses every agent in `active` needs from its neighbours: X (and the
    auxiliary Y under acceleration).  `agents` maps agent id -> object with pack_for(b),
    send_buf[b], send_buf_aux[b], recv_view(b, aux) for the agents this rank owns.  Same-rank
    pairs are device copies; cross-rank pairs are grouped NCCL (or gloo) send/recv.  Every rank
    walks the (active agent, neighbour) pairs in the same order, so sends and receives match."""
    import torch.distributed as dist
    ops = []
    for a in active:
        for b in specs[a].neighbors:
            ob, oa = owner[b], owner[a]
            if ob == rank:
                agents[b].pack_for(a)
            if ob == rank and oa == rank:          # same GPU: device-to-device copy
                agents[a].recv_view(b, False).copy_(agents[b].send_buf[a])
                if acceleration:
                    agents[a].recv_view(b, True).copy_(agents[b].send_buf_aux[a])
            elif ob == rank:                        # send
                ops.append(dist.P2POp(dist.isend, agents[b].send_buf[a], oa))
                if acceleration:
                    ops.append(dist.P2POp(dist.isend, agents[b].send_buf_aux[a], oa))
            elif oa == rank:                        # receive straight into the slot range of b
                ops.append(dist.P2POp(dist.irecv, agents[a].recv_view(b, False), ob))
                if acceleration:
                    ops.append(dist.P2POp(dist.irecv, agents[a].recv_view(b, True), ob))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_poses_host(agents, specs, owner, rank, active, acceleration, group=None, stats=None):
    """The exchange as a deployment of the reference performs it (one process per robot, public poses as
    messages between processes): the poses leave the sender's device into pinned host memory (D2H), travel
    between ranks as HOST buffers (torch.distributed on CPU tensors: gloo), and enter the receiver's device
    from pinned host memory (H2D).  Same walk over (active agent, neighbour) pairs as exchange_poses.
    stats (dict) accumulates the bytes copied each way."""
    import torch
    import torch.distributed as dist
    auxes = (False, True) if acceleration else (False,)
    staged = []                                        # (a, b, aux, host tensor) of the messages this rank holds
    for a in active:
        for b in specs[a].neighbors:
            if owner[b] != rank:
                continue
            agents[b].pack_for(a)
            for aux in auxes:
                src = agents[b].send_buf_aux[a] if aux else agents[b].send_buf[a]
                host = agents[b].host_buffer(("send", a, aux), src.numel())
                host.copy_(src, non_blocking=True)
                staged.append((a, b, aux, host))
                if stats is not None:
                    stats["d2h"] = stats.get("d2h", 0) + host.numel() * 8
    if staged and staged[0][3].is_pinned():
        torch.cuda.current_stream().synchronize()      # the host buffers are complete
    ops, arrived = [], []
    for a in active:
        for b in specs[a].neighbors:
            ob, oa = owner[b], owner[a]
            for aux in auxes:
                if ob == rank and oa == rank:
                    arrived.append((a, b, aux, agents[b].host_buffer(("send", a, aux), 0)))
                elif ob == rank:
                    ops.append(dist.P2POp(dist.isend, agents[b].host_buffer(("send", a, aux), 0), oa, group))
                elif oa == rank:
                    lo, hi = specs[a].nbr_range[b]
                    host = agents[a].host_buffer(("recv", b, aux), (hi - lo) * agents[a].tile)
                    ops.append(dist.P2POp(dist.irecv, host, ob, group))
                    arrived.append((a, b, aux, host))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for a, b, aux, host in arrived:
        agents[a].recv_view(b, aux).copy_(host, non_blocking=True)
        if stats is not None:
            stats["h2d"] = stats.get("h2d", 0) + host.numel() * 8


class NativeExchange:
    """The public-pose exchange inside the C-ABI (dpgo_exchange, csrc/exchange.cu): per round one call that packs
    on the device, moves the cross-rank messages with one NCCL group on the rank's stream and gathers the
    same-device ones straight into the receiver's buffer.  The message list of an active set is built once.
    The communicator is the library's own (ncclCommInitRank from an id rank 0 creates and torch.distributed
    only broadcasts)."""

    def __init__(self, team, device, stream):
        import ctypes as C
        from ._lib import COMM, COMM_ID_BYTES, check, lib
        self.team, self._plans = team, {}
        ident = None
        if team.world > 1:
            import torch.distributed as dist
            box = [None]
            if team.rank == 0:
                buf = C.create_string_buffer(COMM_ID_BYTES)
                check(lib.dpgo_comm_unique_id(buf))
                box[0] = buf.raw
            dist.broadcast_object_list(box, src=0)
            ident = box[0]
        self._comm = COMM()
        check(lib.dpgo_comm_create(int(device), team.rank, team.world, ident, C.c_void_p(stream) if stream else None,
                                   C.byref(self._comm)))
        for ag in team.agents.values():
            ag.native = True
            ag.prob.neighbor_buffer(0)
            ag.prob.neighbor_buffer(1)

    def plan(self, active):
        from ._lib import Message
        key = (tuple(active), self.team.acceleration)
        if key in self._plans:
            return self._plans[key]
        t = self.team
        msgs = []
        for a in active:
            for b in t.specs[a].neighbors:
                src = t.agents[b] if t.owner[b] == t.rank else None
                dst = t.agents[a] if t.owner[a] == t.rank else None
                if src is None and dst is None:
                    continue
                lo, hi = t.specs[a].nbr_range[b]
                for aux in ((0, 1) if t.acceleration else (0,)):
                    m = Message()
                    m.src = src.prob._h.value if src is not None else None
                    m.dst = dst.prob._h.value if dst is not None else None
                    m.peer = t.owner[a] if dst is None else t.owner[b]
                    m.slot = SLOT_Y if aux else SLOT_X
                    m.aux = aux
                    m.count = hi - lo
                    m.d_frames = src.send_idx[a].data_ptr() if src is not None else None
                    m.dst_offset = lo
                    msgs.append(m)
        arr = (Message * max(len(msgs), 1))(*msgs)
        self._plans[key] = (arr, len(msgs))
        return self._plans[key]

    def exchange(self, active):
        from ._lib import check, lib
        arr, n = self.plan(active)
        if n:
            check(lib.dpgo_exchange(self._comm, arr, n))

    def launch_count(self):
        import ctypes as C
        from ._lib import check, lib
        v = C.c_int64()
        check(lib.dpgo_comm_launch_count(self._comm, C.byref(v)))
        return v.value

    def close(self):
        from ._lib import lib
        if self._comm:
            lib.dpgo_comm_destroy(self._comm)
            self._comm = None


class PeerMailboxes:
    """Asynchronous publication of the public poses through peer memory (dpgo_mailbox_* / dpgo_publish / dpgo_collect,
    csrc/exchange.cu): every agent owns one mailbox per neighbour in ITS GPU's memory; the neighbour maps it (CUDA IPC
    across processes, directly inside one process) and stores its poses there after each of its solves; the owner takes a
    consistent snapshot before each of its own solves.  No rendezvous: the ranks run at their own rates, which is the
    reference's asynchronous mode (src/PGOAgent.cpp:475-499, no acceleration)."""

    def __init__(self, team, device):
        import ctypes as C
        from ._lib import IPC_HANDLE_BYTES, MAILBOX, check, lib
        if team.acceleration:
            raise ValueError("the asynchronous mode does not allow acceleration (src/PGOAgent.cpp:477)")
        self.team, self.inbox, self.outbox = team, {}, {}
        mine = {}
        for a, ag in team.agents.items():
            ag.native = True
            ag.prob.neighbor_buffer(0)
            for b in ag.spec.neighbors:
                lo, hi = ag.spec.nbr_range[b]
                buf = C.create_string_buffer(IPC_HANDLE_BYTES)
                mb = MAILBOX()
                check(lib.dpgo_mailbox_create(ag.prob._h, lo, hi - lo, C.byref(mb), buf))
                self.inbox[(a, b)] = mb
                mine[(a, b)] = buf.raw
        table = [mine]
        if team.world > 1:
            import torch.distributed as dist
            table = [None] * team.world
            dist.all_gather_object(table, mine)
        for b, ag in team.agents.items():          # b publishes to every agent a that lists it as a neighbour
            for a in ag.spec.neighbors:
                lo, hi = team.specs[a].nbr_range[b]
                to = MAILBOX()
                if team.owner[a] == team.rank:
                    check(lib.dpgo_mailbox_open(int(device), None, self.inbox[(a, b)], hi - lo, ag.tile, C.byref(to)))
                else:
                    check(lib.dpgo_mailbox_open(int(device), table[team.owner[a]][(a, b)], None, hi - lo, ag.tile,
                                                C.byref(to)))
                self.outbox[(b, a)] = to

    def publish(self, b):
        from ._lib import check, lib
        ag = self.team.agents[b]
        for a in ag.spec.neighbors:
            check(lib.dpgo_publish(ag.prob._h, SLOT_X, self.outbox[(b, a)], ag.send_idx[a].data_ptr()))

    def collect(self, a):
        from ._lib import check, lib
        check(lib.dpgo_collect(self.team.agents[a].prob._h))

    def close(self):
        from ._lib import lib
        for mb in self.outbox.values():
            lib.dpgo_mailbox_close(mb)
        for mb in self.inbox.values():
            lib.dpgo_mailbox_close(mb)
        self.outbox, self.inbox = {}, {}


def block_owner(num_robots, world):
    """Block distribution of agents over ranks (keeps both colours of a chain on every rank)."""
    per_rank = (num_robots + world - 1) // world
    return [min(a // per_rank, world - 1) for a in range(num_robots)]


def build_specs(p1, p2, R, t, kappa, tau, n, d, num_robots):
    p1 = np.asarray(p1, dtype=np.int64); p2 = np.asarray(p2, dtype=np.int64)
    R = np.asarray(R, dtype=np.float64); t = np.asarray(t, dtype=np.float64)
    kappa = np.asarray(kappa, dtype=np.float64); tau = np.asarray(tau, dtype=np.float64)
    ranges, rob, loc = partition(p1, p2, n, num_robots)
    return ranges, [AgentSpec(a, d, ranges, rob, loc, p1, p2, R, t, kappa, tau) for a in range(num_robots)]


class DeviceTeam:
    """All agents of one pose graph, sharded over the ranks of a torch.distributed world
    (block distribution: agent a lives on rank a // (A / world))."""

    def __init__(self, p1, p2, R, t, kappa, tau, n, d, r, num_robots, device=0, stream=None,
                 rank=0, world=1, acceleration=True, params=None, restart_interval=30, native_exchange=False,
                 host_exchange=False, peer_mailboxes=False):
        self.d, self.r, self.n, self.A = d, r, n, num_robots
        self.rank, self.world = rank, world
        self.acceleration = acceleration
        self.ranges, self.specs = build_specs(p1, p2, R, t, kappa, tau, n, d, num_robots)
        self.owner = block_owner(num_robots, world)
        self.colors = color_robot_graph(self.specs)
        if stream is None:
            # all agents, the torch copies and the NCCL ops must be ordered on ONE stream: use
            # torch's current stream (the legacy default stream has the CUDA handle 0x1)
            import torch
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream().cuda_stream or 1
        self.agents = {a: DeviceAgent(self.specs[a], d, r, num_robots, device, stream, acceleration,
                                      restart_interval, params)
                       for a in range(num_robots) if self.owner[a] == rank}
        for a, ag in self.agents.items():
            for b in ag.spec.neighbors:
                # what b needs from a = the frames of a listed in b's neighbour slots for robot a
                ag.prepare_send(b, self.specs[b].nbr_frames[a])
        self.round = 0
        # native_exchange: pack / NCCL send-recv / gather inside the C-ABI on the rank's stream (GPU runs);
        # otherwise torch.distributed P2P ops issued from here (also what the gloo CPU tests drive)
        self.native = NativeExchange(self, device, stream) if native_exchange else None
        # peer_mailboxes: asynchronous publication through peer memory (step_async / publish_all)
        self.mail = PeerMailboxes(self, device) if peer_mailboxes else None
        # host_exchange: the poses travel through pinned host memory and a CPU process group (the end-to-end
        # form: what the reference-facing host API does with them); bytes copied are counted in host_stats
        self.host_exchange = bool(host_exchange)
        self.host_stats = {"d2h": 0, "h2d": 0}
        self.host_group = None
        if self.host_exchange and world > 1:
            import torch.distributed as dist
            self.host_group = dist.new_group(backend="gloo")

    def set_async(self, on):
        """Stream-ordered rounds: no host wait inside a round (see DeviceAgent.async_solve)."""
        for ag in self.agents.values():
            ag.async_solve = bool(on)

    def set_X(self, X):
        dh = self.d + 1
        for a, ag in self.agents.items():
            s, e = self.ranges[a]
            ag.set_X(X[:, s * dh:e * dh])

    # -- public-pose exchange towards the agents in `active` (ref: MultiRobotExample.cpp:183-204)
    def exchange(self, active):
        if self.native is not None:
            self.native.exchange(active)
        elif self.host_exchange:
            exchange_poses_host(self.agents, self.specs, self.owner, self.rank, active, self.acceleration,
                                self.host_group, self.host_stats)
        else:
            exchange_poses(self.agents, self.specs, self.owner, self.rank, active, self.acceleration)

    def step_colored(self):
        """One round of the coloured parallel schedule: the agents of the current colour optimize,
        everybody else performs the non-optimizing iterate (Nesterov bookkeeping)."""
        active = self.colors[self.round % len(self.colors)]
        for a, ag in self.agents.items():
            if a not in active:
                ag.iterate(False)
        self.exchange(active)
        for a, ag in self.agents.items():
            if a in active:
                ag.iterate(True)
        self.round += 1
        return active

    def step_single(self, selected):
        """One iteration of the reference's greedy driver: only `selected` optimizes."""
        for a, ag in self.agents.items():
            if a != selected:
                ag.iterate(False)
        self.exchange([selected])
        if selected in self.agents:
            self.agents[selected].iterate(True)
        self.round += 1

    def step_all(self):
        """Every agent optimizes in the same round from the poses its neighbours published at the
        end of the previous one: the deterministic instance of the asynchronous parallel mode
        (src/PGOAgent.cpp:486-499; no acceleration there, :477).  All GPUs are busy every round."""
        if self.acceleration:
            raise ValueError("the asynchronous / all-agents schedule does not allow acceleration")
        everyone = list(range(self.A))
        self.exchange(everyone)
        for ag in self.agents.values():
            ag.iterate(True)
        self.round += 1
        return everyone

    def publish_all(self):
        """Every local agent stores its public poses into its neighbours' mailboxes (start of an asynchronous run, and
        after every solve)."""
        for b in self.agents:
            self.mail.publish(b)

    def step_async(self):
        """One asynchronous iteration of every LOCAL agent, no coordination with the other ranks: snapshot of whatever
        the neighbours have published so far, local solve, publication (PGOAgent::runOptimizationLoop,
        src/PGOAgent.cpp:486-499, without the Poisson sleep)."""
        if self.mail is None:
            raise ValueError("DeviceTeam(peer_mailboxes=True) is required")
        for a, ag in self.agents.items():
            self.mail.collect(a)
            ag.iterate(True)
            self.mail.publish(a)
        self.round += 1
        return list(self.agents)

    def step_async_lockstep(self, barrier):
        """The same three operations with every rank in step (`barrier()` = device sync + process barrier): all
        snapshots, then all solves, then all publications -- which is exactly step_all, message for message.  Used to
        check the mailboxes against the NCCL exchange."""
        for a in self.agents:
            self.mail.collect(a)
        barrier()
        for ag in self.agents.values():
            ag.iterate(True)
        self.publish_all()
        barrier()
        self.round += 1
        return list(range(self.A))

    def update_weights(self, robust=None):
        """All agents refresh their neighbours' X and re-weight their loop closures (one RobustCost
        per agent, as every PGOAgent owns one).  Returns {agent: (w_private, w_shared)} for the
        agents of this rank."""
        from .robust import RobustCost
        if not hasattr(self, "robust"):
            self.robust = {a: (robust() if callable(robust) else RobustCost()) for a in self.agents}
        acc, self.acceleration = self.acceleration, False      # X only
        try:
            self.exchange(list(range(self.A)))
        finally:
            self.acceleration = acc
        for a, ag in self.agents.items():
            ag.update_measurement_weights(self.robust[a])
        return {a: (ag.w_private.copy(), ag.w_shared.copy()) for a, ag in self.agents.items()}

    def greedy_select(self, central, X=None):
        """Next agent of the reference's greedy driver: largest block of the centralized
        Riemannian gradient (examples/MultiRobotExample.cpp:233-247).  `central` is a
        DeviceProblem of the whole graph."""
        X = self.assemble() if X is None else X
        RG = central.RieGrad(X)
        dh = self.d + 1
        norms = [np.linalg.norm(RG[:, s * dh:e * dh]) for (s, e) in self.ranges]
        return int(np.argmax(norms)), float(np.linalg.norm(RG))

    def assemble(self):
        """Centralized X (host).  With world > 1 every rank gets the full matrix (all_gather)."""
        dh = self.d + 1
        X = np.zeros((self.r, dh * self.n))
        for a, ag in self.agents.items():
            s, e = self.ranges[a]
            X[:, s * dh:e * dh] = ag.get_X()
        if self.world > 1:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(np.ascontiguousarray(X)).cuda()
            dist.all_reduce(t)
            X = t.cpu().numpy()
        return X

    def close(self):
        if self.mail is not None:
            self.mail.close()
        if self.native is not None:
            self.native.close()
        for ag in self.agents.values():
            ag.prob.close()
