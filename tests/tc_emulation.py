"""Host re-run of the tensor-core ADMM kernel's arithmetic (csrc/qp_admm_tc.cu) from the tables the library exports
(``carmpc_qp_get_setup`` selectors 30..43).  Test infrastructure: decodes the swizzled TF32 hi / lo chunk images back into
dense operators, runs the shifted iteration (w^ = w - h, constant columns e) the kernel runs, and returns the state in
the FFMA kernel's terms so that the two formulations can be compared on a machine without a GPU."""
import numpy as np


def sw128_off(r, k):
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4


def tf32_trunc(a):
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


class TcTables:
    def __init__(self, bq):
        g = bq.setup(30).astype(np.int64)
        self.ok, self.np_, self.mp, self.resident, self.na, self.nb, self.b_stage, self.smem = (int(v) for v in g[:8])
        prod = g[8:].reshape(3, 5)
        self.off, self.pair, self.nchunks, self.ksteps, self.ncols = (prod[:, i].astype(int) for i in range(5))
        if not self.ok:
            return
        img = bq.setup(31).astype(np.float32)
        self.B_hi, self.B_lo = [], []
        for p in range(3):
            N, K = int(self.ncols[p]), int(self.ksteps[p]) * 8
            hi = np.zeros((N, K), np.float32); lo = np.zeros((N, K), np.float32)
            r = np.arange(N)[:, None]
            for c in range(int(self.nchunks[p])):
                kk = np.arange(min(32, K - 32 * c))[None, :]
                base = (self.off[p] + c * self.pair[p]) // 4
                o = sw128_off(r, kk) // 4
                hi[:, 32 * c:32 * c + kk.shape[1]] = img[base + o]
                lo[:, 32 * c:32 * c + kk.shape[1]] = img[base + N * 32 + o]
            self.B_hi.append(hi); self.B_lo.append(lo)
        f32 = np.float32
        self.nwd = bq.setup(32).astype(f32)
        self.einv_g = bq.setup(33).astype(f32)
        self.his = bq.setup(34); self.gxs = bq.setup(35).reshape(self.mp, 4); self.gcs = bq.setup(36)
        self.row_id = bq.setup(37).astype(int)
        self.lam = bq.setup(38).astype(f32); self.lb = bq.setup(39).astype(f32); self.ub = bq.setup(40).astype(f32)
        self.einv_b = bq.setup(41).astype(f32); self.nrl = bq.setup(42).astype(f32)
        self.kfv = bq.setup(43).reshape(self.np_, 4)
        self.alpha = f32(bq.opts.alpha)
        self.n, self.m = bq.pq.n, bq.pq.m


def split3(x):
    a = tf32_trunc(x.astype(np.float32))
    r1 = x - a.astype(np.float64)
    b = tf32_trunc(r1.astype(np.float32))
    c = tf32_trunc((r1 - b.astype(np.float64)).astype(np.float32))
    return a, b, c


def mma3(A, Bh, Bl):
    """3xTF32 product as the kernel issues it: A_hi B_hi + A_lo B_hi + A_hi B_lo, float32 accumulate."""
    Ah = tf32_trunc(A)
    Al = tf32_trunc(A - Ah)
    return (Ah @ Bh.T + Al @ Bh.T + Ah @ Bl.T).astype(np.float32)


def emulate_tc(T: TcTables, x0, xref, iters, cd=None):
    """`iters` iterations for states x0 (B, 4).  Returns (w_g (B, m) in logical row order incl. the shift h, w_b (B, n),
    x~ (B, n) of the last iteration, sign (B, m + n))."""
    f32 = np.float32
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    Bn = len(x0)
    cd = np.zeros(Bn) if cd is None else np.asarray(cd, dtype=float)
    NP, mp = T.np_, T.mp
    live = T.row_id >= 0
    h = T.his[None, :] - x0 @ T.gxs.T - cd[:, None] * T.gcs[None, :]
    wg = np.where(live[None, :], -h, 0.0).astype(f32)            # w^ of a cold start (w = 0)
    wb = np.zeros((Bn, NP), f32)
    e = np.zeros((Bn, 16), f32)
    for c in range(4):
        e[:, 3 * c], e[:, 3 * c + 1], e[:, 3 * c + 2] = split3(x0[:, c])
    e[:, 12] = 1.0
    e[:, 13], e[:, 14], e[:, 15] = split3(cd)
    t = (-(T.kfv @ np.asarray(xref, dtype=float))).astype(f32)
    x = np.zeros((Bn, NP), f32)
    for _ in range(iters):
        vb = 2 * np.clip(wb, T.lb, T.ub) - wb
        with np.errstate(invalid="ignore"):
            vg = 2 * np.clip(wg, T.nwd, 0) - wg
        x = mma3(np.concatenate([vb, e, vg], axis=1).astype(f32), T.B_hi[0], T.B_lo[0]) + t[None, :]
        z = T.lam[None, :] * x
        c0 = np.clip(wb, T.lb, T.ub)
        wb = (wb + T.alpha * (z - c0)).astype(f32)
        zh = mma3(np.concatenate([x, e], axis=1).astype(f32), T.B_hi[1], T.B_lo[1])
        with np.errstate(invalid="ignore"):
            c0 = np.clip(wg, T.nwd, 0)
        wg = (wg + T.alpha * (zh - c0)).astype(f32)
    w_log = np.zeros((Bn, T.m), f32)
    w_log[:, T.row_id[live]] = (wg.astype(np.float64) + h)[:, live].astype(f32)
    sign = np.zeros((Bn, T.m + T.n), np.int8)
    sign[:, T.row_id[live]] = ((wg > 0).astype(np.int8) - (wg < T.nwd).astype(np.int8))[:, live]
    sign[:, T.m:] = ((wb > T.ub).astype(np.int8) - (wb < T.lb).astype(np.int8))[:, :T.n]
    return w_log, wb[:, :T.n], x[:, :T.n], sign
