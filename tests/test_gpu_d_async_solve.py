"""GPU test of the stream-ordered local solve (dpgo_optimize_slot_async / dpgo_optimize_result):
the same fused kernel as dpgo_optimize_slot, queued without a host wait.  Results must be
identical to the blocking call, on one handle and over rounds of the multi-agent schedule.
Kept in its own file, last in collection order."""
import numpy as np
import pytest

from oracle import pgo

pytestmark = pytest.mark.gpu


def test_async_solve_equals_blocking_solve(datasets):
    import dpgo_b200
    meas, n, z = datasets("smallGrid3D")
    d, r = meas.d, 5
    X0 = np.asfortranarray(pgo.lifting_matrix(d, r) @ z["T_chordal"])
    gp = dpgo_b200.problem_from_measurements(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r)
    prm = dpgo_b200.default_params()
    gp.slot_set(dpgo_b200.SLOT_Y, X0)
    ref = gp.optimize_slot(dpgo_b200.SLOT_Y, prm)
    X_ref = gp.slot_get(dpgo_b200.SLOT_X)
    # nothing pending yet: the result call must refuse, not return stale data
    with pytest.raises(dpgo_b200.DpgoError):
        gp.optimize_result()
    gp.slot_set(dpgo_b200.SLOT_Y, X0)
    gp.optimize_slot_async(dpgo_b200.SLOT_Y, prm)
    gp.optimize_slot_async(dpgo_b200.SLOT_Y, prm)      # queued twice: the result is the last one's
    res = gp.optimize_result()
    X_async = gp.slot_get(dpgo_b200.SLOT_X)
    assert np.array_equal(X_async, X_ref)
    for k in ("f_init", "f_opt", "gradnorm_init", "gradnorm_opt", "outer_iters", "inner_iters", "accepted",
              "rejected", "tcg_status", "n_qx", "n_precon"):
        assert res[k] == ref[k], k
    assert res["n_launches"] == 1 and res["elapsed_ms"] > 0
    with pytest.raises(dpgo_b200.DpgoError):
        gp.optimize_result()
    # the host-driven solver has host round trips by construction: refused, not silently blocking
    with pytest.raises(dpgo_b200.DpgoError):
        gp.optimize_slot_async(dpgo_b200.SLOT_Y, dpgo_b200.default_params(fused=0))
    gp.close()


@pytest.mark.parametrize("acceleration", [True, False])
def test_stream_ordered_rounds_equal_blocking_rounds(datasets, acceleration):
    from dpgo_b200 import rbcd
    meas, n, z = datasets("smallGrid3D")
    d, r, A = meas.d, 5, 5
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    out = {}
    for mode in (False, True):
        team = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                               acceleration=acceleration)
        team.set_async(mode)
        team.set_X(X0)
        for _ in range(8):
            team.step_colored()
        res = {a: ag.result() for a, ag in team.agents.items()}
        out[mode] = (team.assemble(), res)
        team.close()
    assert np.array_equal(out[True][0], out[False][0])
    for a in out[False][1]:
        assert out[True][1][a]["f_opt"] == out[False][1][a]["f_opt"]
        assert out[True][1][a]["inner_iters"] == out[False][1][a]["inner_iters"]


@pytest.mark.parametrize("schedule", ["colored", "all"])
def test_native_exchange_equals_python_exchange(datasets, schedule):
    """dpgo_exchange (pack / gather inside the C-ABI, csrc/exchange.cu) against the exchange issued from Python
    (the path the oracle-parity tests of the team use): identical poses after rounds that cross the Nesterov
    restart, on one device (same-device messages; the NCCL messages are covered by tests/gpu_native_exchange_2gpu.py
    under torchrun)."""
    from dpgo_b200 import rbcd
    meas, n, z = datasets("smallGrid3D")
    d, r, A = meas.d, 5, 5
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    out = {}
    for native in (False, True):
        team = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A,
                               acceleration=(schedule == "colored"), native_exchange=native)
        team.set_async(True)
        team.set_X(X0)
        for _ in range(33 if schedule == "colored" else 6):
            (team.step_colored if schedule == "colored" else team.step_all)()
        out[native] = team.assemble()
        if native:
            assert team.native.launch_count() > 0
        team.close()
    assert np.array_equal(out[True], out[False])


def test_native_exchange_across_ranks():
    """Two ranks, two GPUs (skipped on a one-GPU box): the NCCL messages of dpgo_exchange give the same poses as
    one rank holding all agents."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(root, "tests", "gpu_native_exchange_2gpu.py")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "native exchange ok" in out.stdout


def test_peer_mailboxes_lockstep_equals_all_agents_schedule(datasets):
    """The asynchronous publication (mailboxes in the receiver's memory, dpgo_publish / dpgo_collect) run in lockstep
    is the all-agents schedule message for message: identical poses after 6 rounds on torus3D / 8 agents (BASELINE
    configs[3]); free-running on one device (agents one after another, each seeing its predecessors' new poses) it is
    a block Gauss-Seidel sweep: finite, below the initial cost and within 1 % of the all-agents cost after the same
    number of solves."""
    import torch
    from dpgo_b200 import rbcd
    meas, n, z = datasets("torus3D")
    d, r, A = meas.d, 5, 8
    X0 = pgo.lifting_matrix(d, r) @ z["T_chordal"]
    central = pgo.QuadraticProblem(pgo.connection_laplacian(meas, n), np.zeros((r, (d + 1) * n)), d)

    def make(**kw):
        t = rbcd.DeviceTeam(meas.p1, meas.p2, meas.R, meas.t, meas.kappa, meas.tau, n, d, r, A, acceleration=False, **kw)
        t.set_async(True)
        t.set_X(X0)
        return t

    ref = make(native_exchange=True)
    for _ in range(6):
        ref.step_all()
    X_all = ref.assemble()
    ref.close()
    lock = make(peer_mailboxes=True)
    lock.publish_all()
    for _ in range(6):
        lock.step_async_lockstep(torch.cuda.synchronize)
    X_lock = lock.assemble()
    lock.close()
    assert np.array_equal(X_lock, X_all)
    free = make(peer_mailboxes=True)
    free.publish_all()
    for _ in range(6):
        free.step_async()
    X_free = free.assemble()
    free.close()
    assert np.isfinite(X_free).all()
    assert central.f(X_free) < central.f(X0)
    assert abs(central.f(X_free) - central.f(X_all)) <= 1e-2 * central.f(X_all)
