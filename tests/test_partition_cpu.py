"""CPU tests of the host logic behind the two-level preconditioner: the nested dissection of the
pose graph (dpgo_b200/csrc/dissect.h through the host-only C-ABI entry dpgo_two_level_partition).
The algebra built on it is exact for ANY partition with this separator property, which is checked
here; that the device result equals the reference's (Q + 0.1 I)^-1 solve is a GPU test."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import pgo


def _partition(n, rowptr, colidx, dh, max_poses=0):
    from dpgo_b200 import _lib
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    group = np.empty(max(n, 1), dtype=np.int32)
    k = C.c_int()
    ip = C.POINTER(C.c_int32)
    rc = _lib.lib.dpgo_two_level_partition(n, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), dh, max_poses,
                                           group.ctypes.data_as(ip), C.byref(k))
    assert rc == 0, _lib.lib.dpgo_last_error()
    return group[:n], k.value


def _pose_graph(meas, n):
    A = sp.coo_matrix((np.ones(len(meas)), (meas.p1, meas.p2)), shape=(n, n))
    A = (A + A.T + sp.identity(n)).tocsr()
    A.sort_indices()
    return A


def _check(A, group, K, thr):
    n = A.shape[0]
    assert group.min() >= -1 and group.max() == K - 1
    sizes = np.bincount(group[group >= 0], minlength=K)
    assert sizes.min() >= 1                                  # no empty domain
    coo = A.tocoo()
    gi, gj = group[coo.row], group[coo.col]
    both = (gi >= 0) & (gj >= 0)
    assert np.all(gi[both] == gj[both])                      # no edge joins two different domains
    return sizes, int(np.sum(group < 0))


@pytest.mark.parametrize("name", ["smallGrid3D", "sphere2500", "torus3D", "grid3D", "city10000"])
def test_partition_properties_on_reference_datasets(datasets, name):
    meas, n, _ = datasets(name)
    dh = meas.d + 1
    A = _pose_graph(meas, n)
    thr = max(8, 320 // dh)                                  # kDdStages * kStageK / dh
    group, K = _partition(n, A.indptr, A.indices, dh)
    sizes, nsep = _check(A, group, K, thr)
    assert sizes.max() <= thr                                # these graphs always split
    assert K >= max(1, n // thr)
    # the separator is what makes the Schur complement dense: keep it a modest part of the graph
    assert nsep <= 0.45 * n, (name, nsep, n)
    g2, K2 = _partition(n, A.indptr, A.indices, dh)
    assert K2 == K and np.array_equal(g2, group)             # deterministic


def test_partition_edge_cases():
    # a single pose, a graph below the threshold, a chain, a star (cannot be split), two components
    g, K = _partition(1, [0, 1], [0], 4)
    assert K == 1 and g[0] == 0
    chain = sp.diags([np.ones(49), np.ones(50), np.ones(49)], [-1, 0, 1]).tocsr()
    g, K = _partition(50, chain.indptr, chain.indices, 4)
    assert K == 1 and np.all(g == 0)                         # 50 <= 80 poses: one domain, no separator
    g, K = _partition(50, chain.indptr, chain.indices, 4, max_poses=8)
    sizes, nsep = _check(chain, g, K, 8)
    assert sizes.max() <= 8 and 4 <= nsep <= 12              # single-pose separators along the chain
    m = 40
    star = sp.lil_matrix((m, m)); star[0, :] = 1; star[:, 0] = 1; star.setdiag(1)
    star = star.tocsr()
    g, K = _partition(m, star.indptr, star.indices, 4, max_poses=8)
    _check(star, g, K, 8)                                    # BFS depth 1-2: separator = the hub or nothing
    two = sp.block_diag([chain, chain]).tocsr()
    g, K = _partition(100, two.indptr, two.indices, 4, max_poses=60)
    sizes, nsep = _check(two, g, K, 60)
    assert K == 2 and nsep == 0 and sorted(sizes) == [50, 50]  # components are split off without a separator


def test_partition_rejects_bad_input():
    from dpgo_b200 import _lib
    ip = C.POINTER(C.c_int32)
    rowptr = np.array([0, 1, 2], dtype=np.int32)
    colidx = np.array([0, 7], dtype=np.int32)                # column out of range
    group = np.zeros(2, dtype=np.int32)
    k = C.c_int()
    rc = _lib.lib.dpgo_two_level_partition(2, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), 4, 0,
                                           group.ctypes.data_as(ip), C.byref(k))
    assert rc == -1


def test_two_level_elimination_is_exact_for_this_partition(datasets):
    """The algebra of the two-level variant on the host (dense numpy): with A = Q + 0.1 I permuted to
    [domains | separator], z_S = S^-1 (r_S - A_SI A_II^-1 r_I), z_I = A_II^-1 (r_I - A_IS z_S)
    equals A^-1 r, and A_II is block diagonal over the domains."""
    meas, n, _ = datasets("smallGrid3D")
    d, dh = meas.d, meas.d + 1
    Q = pgo.connection_laplacian(meas, n)
    A = (Q + 0.1 * sp.identity(dh * n)).toarray()
    G = _pose_graph(meas, n)
    group, K = _partition(n, G.indptr, G.indices, dh, max_poses=20)
    assert K >= 4 and np.sum(group < 0) > 0
    sc = lambda poses: (np.asarray(poses)[:, None] * dh + np.arange(dh)).ravel()
    I = np.concatenate([sc(np.where(group == k)[0]) for k in range(K)])
    S = sc(np.where(group < 0)[0])
    AII, AIS, ASI, ASS = A[np.ix_(I, I)], A[np.ix_(I, S)], A[np.ix_(S, I)], A[np.ix_(S, S)]
    off = 0
    mask = np.zeros_like(AII, dtype=bool)
    for k in range(K):
        m = dh * int(np.sum(group == k))
        mask[off:off + m, off:off + m] = True
        off += m
    assert np.all(AII[~mask] == 0)                           # interior blocks decouple
    rng = np.random.default_rng(0)
    r = rng.standard_normal(dh * n)
    AIIinv = np.linalg.inv(AII)
    Sig = ASS - ASI @ AIIinv @ AIS
    zS = np.linalg.solve(Sig, r[S] - ASI @ (AIIinv @ r[I]))
    zI = AIIinv @ (r[I] - AIS @ zS)
    z = np.zeros(dh * n); z[I] = zI; z[S] = zS
    ref = np.linalg.solve(A, r)
    assert np.linalg.norm(z - ref) <= 1e-10 * np.linalg.norm(ref)
