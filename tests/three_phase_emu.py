"""numpy replay of the three-phase form of the two-level preconditioner from the index plan the
library builds (dpgo_three_phase_plan, dpgo_b200/csrc/dd_plan.h): stage buffers are filled from
numpy inverses exactly as the device set-up lays them out, the three strip phases and the finish
are replayed with the staging rules of the CUDA kernels (phase_strip_gemv, SRC modes 1 / 2 / 3),
and the result is compared with the exact (Q + 0.1 I)^-1 by the caller.  Test infrastructure."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

STAGE_K, COLS = 32, 64
SECTIONS = ["scalars", "group", "pcol", "srow", "icol", "dom_off", "dom_m", "dom_pad", "t_off", "t_m", "t_pad",
            "sk_ptr", "sk", "tptr", "tcol", "gchunk", "gidx", "strips1", "strips3", "strips5", "cta1", "cta3", "cta5"]
SCALARS = ["n", "dh", "K", "nS", "V", "sep_col0", "pcols", "ycols", "nsplit3", "stages1", "stages3", "stages5",
           "bytes_per_apply"]


def fetch_plan(fn, n, rowptr, colidx, dh, max_poses=0, V=148, split=0, affine=0):
    """fn = the C entry point (dpgo_three_phase_plan signature)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    ip, lp = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    need = C.c_int64()
    rc = fn(n, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), dh, max_poses, V, split, affine, None, 0,
            C.byref(need))
    assert rc == 0
    img = np.zeros(need.value, dtype=np.int64)
    rc = fn(n, rowptr.ctypes.data_as(ip), colidx.ctypes.data_as(ip), dh, max_poses, V, split, affine,
            img.ctypes.data_as(lp), img.size, C.byref(need))
    assert rc == 0 and need.value == img.size
    ns = int(img[0])
    assert ns == len(SECTIONS)
    plan = {}
    for s, name in enumerate(SECTIONS):
        off, ln = int(img[1 + 2 * s]), int(img[2 + 2 * s])
        plan[name] = img[off:off + ln].copy()
    for k, name in enumerate(SCALARS):
        plan[name] = int(plan["scalars"][k])
    for ph in ("1", "3", "5"):
        plan["strips" + ph] = plan["strips" + ph].reshape(-1, 8)
    return plan


def dense_blocks(A, plan):
    """M_k = A_k^-1, C_k = M_k A_kS restricted to the columns of S_k, Sigma^-1 (all dense numpy)."""
    dh, K, nS = plan["dh"], plan["K"], plan["nS"]
    group, srow = plan["group"], plan["srow"]
    sc = lambda poses: (np.asarray(poses, dtype=np.int64)[:, None] * dh + np.arange(dh)).ravel()
    A = sp.csr_matrix(A)
    doms = [np.where(group == k)[0] for k in range(K)]          # ascending pose ids = the plan's order
    S = sc(srow)
    ASS = A[S][:, S].toarray() if nS else np.zeros((0, 0))
    M, Cc = [], []
    Sig = ASS.copy()
    for k in range(K):
        I = sc(doms[k])
        Mk = np.linalg.inv(A[I][:, I].toarray())
        M.append(Mk)
        sk = plan["sk"][plan["sk_ptr"][k]:plan["sk_ptr"][k + 1]]
        if len(sk):
            cols = sc(sk)                                       # scalar columns inside the S order
            AkS = A[I][:, S[cols]].toarray()
            Ck = Mk @ AkS
            Sig[np.ix_(cols, cols)] -= AkS.T @ Ck
            Cc.append(Ck)
        else:
            Cc.append(np.zeros((len(I), 0)))
        # the plan's S_k must be exactly the separator poses coupled to the domain
        if nS:
            full = A[I][:, S].toarray().reshape(len(I), nS, dh)
            touched = np.where(np.abs(full).sum(axis=(0, 2)) > 0)[0]
            assert np.array_equal(touched, sk), (k, touched, sk)
    SigInv = np.linalg.inv(Sig) if nS else Sig
    return M, Cc, SigInv


def fill_stage_buffers(plan, M, Cc, SigInv):
    """What the device set-up writes (k_dd_layout / k_dd_layout_rect): stage (kk, jj) tiles."""
    def padded(X, rows, cols):
        P = np.zeros((rows, cols))
        P[:X.shape[0], :X.shape[1]] = X
        return P
    bufs = {ph: np.full((plan["stages" + ph], STAGE_K, COLS), np.nan) for ph in ("1", "3", "5")}
    sep_chunk0 = plan["sep_col0"] // STAGE_K
    for ph in ("1", "3", "5"):
        for cb, kc0, nch, slot, off, kind, k, blk in plan["strips" + ph]:
            for c in range(nch):
                if kind == 0:       # M_k: out column 64 blk + jj, inner 32 c + kk
                    P = padded(M[k], plan["dom_pad"][k], plan["dom_pad"][k])
                    tile = P[64 * blk:64 * blk + 64, 32 * c:32 * c + 32].T
                elif kind == 1:     # C_k[:, S_k]: out = compact S_k column 64 blk + jj, inner = domain row 32 c + kk
                    P = padded(Cc[k], plan["dom_pad"][k], plan["t_pad"][k])
                    tile = P[32 * c:32 * c + 32, 64 * blk:64 * blk + 64]
                elif kind == 2:     # Sigma^-1: out column 64 blk + jj, inner chunk (kc0 - sep_chunk0) + c
                    padS = plan["pcols"] - plan["sep_col0"]
                    P = padded(SigInv, padS, padS)
                    ch = kc0 - sep_chunk0 + c
                    tile = P[64 * blk:64 * blk + 64, 32 * ch:32 * ch + 32].T
                else:               # C_k[:, S_k]^T: out = domain row 64 blk + jj, inner = compact column 32 c + kk
                    P = padded(Cc[k], plan["dom_pad"][k], -(-plan["t_m"][k] // STAGE_K) * STAGE_K)
                    tile = P[64 * blk:64 * blk + 64, 32 * c:32 * c + 32].T
                assert np.isnan(bufs[ph][off + c]).all(), "two strips share a stage"
                bufs[ph][off + c] = tile
        assert not np.isnan(bufs[ph]).any(), "stage buffer has unwritten stages"
    return bufs


def run_strips(plan, ph, buf, stage_value, out):
    """out[slot][:, 64 cb + jj] = sum_c sum_kk stage_value(32 (kc0 + c) + kk) * buf[off + c][kk, jj]"""
    written = set()
    for cb, kc0, nch, slot, off, kind, k, blk in plan["strips" + ph]:
        acc = np.zeros((out.shape[1], COLS))
        for c in range(nch):
            v = np.stack([stage_value(STAGE_K * (kc0 + c) + kk) for kk in range(STAGE_K)], axis=1)   # R x 32
            acc += v @ buf[off + c]
        assert (slot, cb) not in written, "two strips write the same (slot, column block)"
        written.add((slot, cb))
        out[slot][:, COLS * cb:COLS * cb + COLS] = acc
    return written


def apply(plan, bufs, r):
    """r: R x N in the original column order.  Returns z = r (Q + 0.1 I)^-1 by the three phases."""
    R = r.shape[0]
    icol, tptr, tcol, gidx = plan["icol"], plan["tptr"], plan["tcol"], plan["gidx"]
    sep0, pcols, ycols, ns3 = plan["sep_col0"], plan["pcols"], plan["ycols"], plan["nsplit3"]
    zero = np.zeros(R)
    y = np.zeros((1, R, ycols))
    zs = np.zeros((ns3, R, pcols))
    w = np.zeros((1, R, pcols))

    def src1(col):                       # SRC 1: gathered through icol
        return r[:, icol[col]] if icol[col] >= 0 else zero

    def src3(col):                       # SRC 2: r_S minus the scattered g terms
        v = src1(col).copy()
        j = col - sep0
        for e in range(tptr[j], tptr[j + 1]):
            v -= y[0][:, tcol[e]]
        return v

    def src5(i):                         # SRC 3: sum of the partial slots of z_S through the gather list
        return zs[:, :, gidx[i]].sum(axis=0) if gidx[i] >= 0 else zero

    run_strips(plan, "1", bufs["1"], src1, y)
    run_strips(plan, "3", bufs["3"], src3, zs)
    run_strips(plan, "5", bufs["5"], src5, w)
    dh = plan["dh"]
    z = np.zeros_like(r)
    for i in range(plan["n"]):
        pc = plan["pcol"][i]
        cols = slice(pc, pc + dh)
        if pc >= sep0:
            z[:, i * dh:(i + 1) * dh] = zs[:, :, cols].sum(axis=0)
        else:
            z[:, i * dh:(i + 1) * dh] = y[0][:, cols] - w[0][:, cols]
    return z


def check_tables(plan):
    """Structural properties the kernels rely on."""
    V = plan["V"]
    for ph in ("1", "3", "5"):
        cta, strips = plan["cta" + ph], plan["strips" + ph]
        assert len(cta) == V + 1 and cta[0] == 0 and cta[-1] == len(strips) and np.all(np.diff(cta) >= 0)
        assert np.all(strips[:, 2] >= (0 if ph == "5" else 1))     # phase 5: empty strips of isolated domains
    assert plan["sep_col0"] % COLS == 0 and plan["pcols"] % COLS == 0 and plan["ycols"] % COLS == 0
    assert np.all(plan["dom_off"] % COLS == 0) and np.all(plan["t_off"] % COLS == 0)
    assert len(plan["gidx"]) % STAGE_K == 0 and len(plan["tptr"]) == plan["pcols"] - plan["sep_col0"] + 1
    g = plan["gidx"]
    assert np.all((g == -1) | ((g >= plan["sep_col0"]) & (g < plan["pcols"])))
    t = plan["tcol"]
    assert np.all((t >= plan["pcols"]) & (t < plan["ycols"])) and len(set(t.tolist())) == len(t)
    # icol is a bijection between the original columns and the non-padding permuted columns
    ic = plan["icol"]
    real = ic[ic >= 0]
    assert np.array_equal(np.sort(real), np.arange(plan["n"] * plan["dh"]))
    assert np.all(ic[plan["pcols"]:] == -1)
