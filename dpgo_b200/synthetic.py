"""Synthetic 3-D grid pose graphs for roofline-scale measurements ("a synthetic 3-D
grid pose graph with >= 1 M poses: odometry chain + 6-neighbour lattice loop closures,
R = I perturbed by N(0, 0.05^2) rad, unit steps, kappa = 200, tau = 100").  Deterministic
(numpy default_rng(seed)); used by tools/qx_scale.py and the scale tests."""
import numpy as np


def _rodrigues(w):
    """exp of a batch of rotation vectors (m, 3) -> (m, 3, 3)."""
    th = np.linalg.norm(w, axis=1)
    k = w / np.maximum(th, 1e-300)[:, None]
    K = np.zeros((len(w), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def grid3d(L, seed=0, rot_sigma=0.05, noise_rot=0.01, noise_t=0.01, kappa=200.0, tau=100.0):
    """L^3 poses visited along a boustrophedon (snake) path so that consecutive pose ids are
    lattice neighbours (the odometry chain); every other lattice edge is a loop closure.
    Returns dict(p1, p2, R, t, kappa, tau, n, d, T_true)."""
    rng = np.random.default_rng(seed)
    n = L ** 3
    # snake ordering: id -> (x, y, z)
    ids = np.arange(n)
    z = ids // (L * L)
    rem = ids % (L * L)
    y = rem // L
    y = np.where(z % 2 == 0, y, L - 1 - y)
    x = rem % L
    row = z * L + (rem // L)
    x = np.where(row % 2 == 0, x, L - 1 - x)
    pos = np.stack([x, y, z], axis=1).astype(np.float64)
    coord_to_id = np.empty((L, L, L), dtype=np.int64)
    coord_to_id[x, y, z] = ids
    Rw = _rodrigues(rot_sigma * rng.standard_normal((n, 3)))
    # lattice edges (+x, +y, +z neighbours), oriented from the smaller to the larger pose id
    tails, heads = [], []
    for ax in range(3):
        sl_a = [slice(None)] * 3
        sl_b = [slice(None)] * 3
        sl_a[ax] = slice(0, L - 1)
        sl_b[ax] = slice(1, L)
        a = coord_to_id[tuple(sl_a)].ravel()
        b = coord_to_id[tuple(sl_b)].ravel()
        tails.append(np.minimum(a, b))
        heads.append(np.maximum(a, b))
    p1 = np.concatenate(tails)
    p2 = np.concatenate(heads)
    # odometry edges first (p2 == p1 + 1), as the reference's datasets list them
    order = np.lexsort((p2, p1, ~(p2 == p1 + 1)))
    p1, p2 = p1[order], p2[order]
    m = len(p1)
    Ri, Rj = Rw[p1], Rw[p2]
    Rn = _rodrigues(noise_rot * rng.standard_normal((m, 3)))
    R = np.einsum("mba,mbc->mac", Ri, Rj) @ Rn                  # R_i^T R_j * noise
    t = np.einsum("mba,mb->ma", Ri, pos[p2] - pos[p1]) + noise_t * rng.standard_normal((m, 3))
    T_true = np.zeros((3, 4 * n))
    T_true[:, np.arange(n) * 4 + 3] = pos.T
    for c in range(3):
        T_true[:, np.arange(n) * 4 + c] = Rw[:, :, c].T
    return dict(p1=p1.astype(np.int32), p2=p2.astype(np.int32), R=R, t=t,
                kappa=np.full(m, kappa), tau=np.full(m, tau), n=n, d=3, T_true=T_true)
