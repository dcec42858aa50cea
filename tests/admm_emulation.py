"""Host re-run of the CUDA ADMM kernel's arithmetic from the padded tables the library exports
(``carmpc_qp_get_setup`` selectors 10..26).  Test infrastructure: validates the host setup (permutation into
independent chains, V order, owner order, K-range segments) on a machine without a GPU, and predicts what the
kernel computes.  Mirrors csrc/qp_admm.cu stage by stage in float32."""
import numpy as np


class KernelTables:
    def __init__(self, bq):
        g = bq.setup(26).astype(int)
        self.n, self.m, self.nA, self.mp, self.npad4, self.mv4, self.ktot, self.nGA, self.nGB = g
        f32 = np.float32
        self.P = bq.setup(10).astype(f32).reshape(self.nA, self.ktot)
        self.Gs = bq.setup(11).astype(f32).reshape(self.mp, self.npad4)
        self.GsT = bq.setup(12).astype(f32).reshape(self.nA, self.mv4)
        self.his = bq.setup(13)
        self.Gxs = bq.setup(14).reshape(self.mp, 4)
        self.width = bq.setup(15).astype(f32)
        self.vpos = bq.setup(16).astype(int)
        self.row_id = bq.setup(17).astype(int)
        self.lam = bq.setup(18).astype(f32)
        self.lbs = bq.setup(19).astype(f32)
        self.ubs = bq.setup(20).astype(f32)
        self.KF = bq.setup(21).reshape(self.nA, 4)
        self.var_id = bq.setup(22).astype(int)
        self.segA = bq.setup(23).astype(int).reshape(self.nGA, 4)
        self.segB = bq.setup(24).astype(int).reshape(self.nGB, 2)
        self.Dsc = bq.setup(25).astype(f32)
        t = bq.tiling()
        self.RA, self.RB = 5, 7
        self.alpha = np.float32(bq.opts.alpha)


def emulate(T: KernelTables, x0, xref, iters, return_trace=False, return_state=False):
    """Runs `iters` iterations for a batch x0 (B, 4).  Returns (u (B, n) unscaled, sign (B, m + n)) and optionally the
    residual trace or the final ADMM state of the general rows in LOGICAL row order (w, lo, hi: (B, m), scaled)."""
    f32 = np.float32
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    B = len(x0)
    dx = x0 - np.asarray(xref, dtype=float)[None, :]
    x0t = (dx @ T.KF.T).astype(f32)                               # (B, nA)
    hi = (T.his[None, :] - x0 @ T.Gxs.T)
    hi = np.where(T.row_id[None, :] >= 0, hi, 3.0e38).astype(f32)   # (B, mp)
    with np.errstate(invalid="ignore", over="ignore"):
        lo = hi - T.width[None, :]
    wB = np.zeros((B, T.mp), dtype=f32)
    wA = np.zeros((B, T.nA), dtype=f32)
    V = np.zeros((B, T.ktot), dtype=f32)
    live = T.row_id >= 0
    vp = T.vpos
    cB = np.clip(wB, lo, hi)
    V[:, vp[live]] = (2 * cB - wB)[:, live]
    cA = np.clip(wA, T.lbs, T.ubs)
    V[:, T.mv4:T.mv4 + T.nA] = np.where(T.var_id[None, :] >= 0, 2 * cA - wA, 0)
    xt = np.zeros((B, T.npad4), dtype=f32)
    res_trace = []
    for it in range(iters):
        # stage A
        for ga in range(T.nGA):
            r0 = ga * T.RA
            gb, ge, bb, be = T.segA[ga]
            acc = x0t[:, r0:r0 + T.RA].copy()
            acc += V[:, gb:ge] @ T.P[r0:r0 + T.RA, gb:ge].T
            acc += V[:, bb:be] @ T.P[r0:r0 + T.RA, bb:be].T
            xt[:, r0:r0 + T.RA] = acc
        # box rows
        z = T.lam[None, :] * xt[:, :T.nA]
        c0 = np.clip(wA, T.lbs, T.ubs)
        wA = wA + T.alpha * (z - c0)
        c1 = np.clip(wA, T.lbs, T.ubs)
        V[:, T.mv4:T.mv4 + T.nA] = 2 * c1 - wA
        # stage B
        zB = np.zeros((B, T.mp), dtype=f32)
        for g in range(T.nGB):
            r0 = g * T.RB
            kb, ke = T.segB[g]
            if ke > kb:
                zB[:, r0:r0 + T.RB] = xt[:, kb:ke] @ T.Gs[r0:r0 + T.RB, kb:ke].T
        c0 = np.clip(wB, lo, hi)
        wB = wB + T.alpha * (zB - c0)
        c1 = np.clip(wB, lo, hi)
        V[:, vp[live]] = (2 * c1 - wB)[:, live]
        if return_trace:
            res_trace.append(np.abs(zB - c1)[:, live].max(1))
    u = np.zeros((B, T.n))
    sign = np.zeros((B, T.m + T.n), dtype=np.int8)
    vid = T.var_id
    ok = vid >= 0
    u[:, vid[ok]] = (T.Dsc[None, :] * xt[:, :T.nA])[:, ok]
    sign[:, T.m + vid[ok]] = ((wA > T.ubs).astype(np.int8) - (wA < T.lbs).astype(np.int8))[:, ok]
    sign[:, T.row_id[live]] = ((wB > hi).astype(np.int8) - (wB < lo).astype(np.int8))[:, live]
    if return_trace:
        return u, sign, np.array(res_trace)
    if return_state:
        w = np.zeros((B, T.m), dtype=f32)
        w[:, T.row_id[live]] = wB[:, live]
        return u, sign, w
    return u, sign


def polish_reference(pq, x0, xref, sign, rounds=8):
    """numpy float64 statement of csrc/qp_polish.cu for one sample.  Returns (u, certified)."""
    n, m = pq.n, pq.m
    A = np.vstack((pq.G, np.eye(n)))
    Hinv = np.linalg.inv(pq.H)
    AH = A @ Hinv
    AHA = AH @ A.T
    q = pq.F @ (x0 - xref)
    hi = np.hstack((pq.hi - pq.Gx @ x0, pq.ub))
    lo = np.hstack((pq.lo - pq.Gx @ x0, pq.lb))
    sgn = sign.astype(int).copy()
    uunc = -Hinv @ q
    Auu = A @ uunc
    u = uunc
    for _ in range(rounds):
        act = np.flatnonzero(sgn)
        if len(act):
            b = np.where(sgn[act] > 0, hi[act], lo[act])
            M = AHA[np.ix_(act, act)]
            M = M + 1e-13 * np.diag(np.diag(M))
            lam = np.linalg.lstsq(M, Auu[act] - b, rcond=1e-11)[0]
            u = uunc - AH[act].T @ lam
        else:
            lam = np.zeros(0)
            u = uunc
        Au = A @ u
        viol = np.maximum(Au - hi, lo - Au)
        viol[~np.isfinite(viol)] = -np.inf
        worst = int(np.argmax(viol))
        bad = act[lam * sgn[act] < -1e-9 * (1 + (np.abs(lam).max() if len(lam) else 0))] if len(act) else []
        if viol[worst] <= 1e-8 and len(bad) == 0:
            return u, True
        sgn[bad] = 0
        if viol[worst] > 1e-8:
            sgn[worst] = 1 if Au[worst] - hi[worst] >= lo[worst] - Au[worst] else -1
    return u, False
