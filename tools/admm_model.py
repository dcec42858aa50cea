"""Numpy model of the device ADMM iteration (development / test aid, not product code).

It mirrors, operation for operation, what ``carmpc_b200/csrc/qp_admm.cu`` does per sample, vectorised over the
batch, in a selectable dtype, so that the algorithm (scaling, rho, relaxation, the single-vector (w) state,
residual and certificate formulas, check cadence) can be tuned and validated on the CPU.  tests/ use it to
cross-check the host setup exported by the C library and as a float32 predictor of kernel results.
"""
from __future__ import annotations

import numpy as np


def ruiz_scale(H, G, n_iter=15):
    """Modified Ruiz equilibration of [[H, A'], [A, 0]] with A = [G; I].  Returns D (n), Eg (m), Eb (n), c."""
    n = H.shape[0]
    m = G.shape[0]
    D = np.ones(n)
    Eg = np.ones(m)
    Eb = np.ones(n)
    c = 1.0
    for _ in range(n_iter):
        Hs = c * (D[:, None] * H * D[None, :])
        Gs = Eg[:, None] * G * D[None, :]
        Bs = Eb * D                                            # diagonal of the scaled box rows
        col = np.maximum.reduce([np.abs(Hs).max(0), np.abs(Gs).max(0) if m else np.zeros(n), np.abs(Bs)])
        rowg = np.abs(Gs).max(1) if m else np.zeros(0)
        rowb = np.abs(Bs)
        col = np.where(col < 1e-4, 1.0, col)
        rowg = np.where(rowg < 1e-4, 1.0, rowg)
        rowb = np.where(rowb < 1e-4, 1.0, rowb)
        D = D / np.sqrt(col)
        Eg = Eg / np.sqrt(rowg)
        Eb = Eb / np.sqrt(rowb)
        # cost normalisation (OSQP): mean column inf-norm of the scaled Hessian
        Hs = c * (D[:, None] * H * D[None, :])
        avg = np.abs(Hs).max(0).mean()
        if avg > 1e-4:
            c = c / avg
    return D, Eg, Eb, c


class DeviceModel:
    def __init__(self, pq, rho=0.1, alpha=1.6, scaling_iters=15, dtype=np.float32, precise=False):
        self.pq = pq
        self.dtype = dtype
        self.precise = precise
        n, m = pq.n, pq.m
        if scaling_iters > 0:
            D, Eg, Eb, c = ruiz_scale(pq.H, pq.G, scaling_iters)
        else:
            D, Eg, Eb, c = np.ones(n), np.ones(m), np.ones(n), 1.0
        self.D, self.Eg, self.Eb, self.c = D, Eg, Eb, c
        self.rho, self.alpha = rho, alpha
        Hs = c * (D[:, None] * pq.H * D[None, :])
        Gs = Eg[:, None] * pq.G * D[None, :]
        lam = Eb * D
        K = Hs + rho * (Gs.T @ Gs + np.diag(lam * lam))
        self.Kinv64 = np.linalg.inv(K)
        self.Hs64, self.Gs64 = Hs, Gs
        f = dtype
        self.Kinv = self.Kinv64.astype(f)
        self.Gs = Gs.astype(f)
        self.lam = lam.astype(f)
        self.Fs = (c * D[:, None] * pq.F)                  # q_s = Fs (x0 - xref)            (fp64 on device)
        self.Gxs = Eg[:, None] * pq.Gx                     # bound_s = his - Gxs x0          (fp64 on device)
        self.his = Eg * pq.hi
        self.width = np.where(np.isfinite(pq.lo), Eg * (pq.hi - pq.lo), np.inf)
        self.ubs = Eb * pq.ub
        self.lbs = Eb * pq.lb
        self.Einv_g = (1.0 / Eg).astype(f)
        self.Einv_b = (1.0 / Eb).astype(f)
        self.Dinv = (1.0 / D).astype(f)

    def solve(self, x0, xref=None, eps_abs=1e-5, eps_rel=1e-5, eps_inf=1e-4, max_iter=4000, check_every=10,
              verbose=False):
        pq, f = self.pq, self.dtype
        x0 = np.atleast_2d(np.asarray(x0, dtype=float))
        xref = pq.goal if xref is None else np.asarray(xref, dtype=float)
        Bn = len(x0)
        n, m = pq.n, pq.m
        # per-sample data, computed in float64 then rounded (as the kernel does)
        q = ((x0 - xref[None, :]) @ self.Fs.T).astype(f)                     # (B, n)
        hi = (self.his[None, :] - x0 @ self.Gxs.T)                           # (B, m) fp64
        lo = (hi - self.width[None, :]).astype(f)
        hi = hi.astype(f)
        ubs, lbs = self.ubs.astype(f), self.lbs.astype(f)
        pre_ok = np.ones(Bn, dtype=bool)
        if len(pq.pre_hi):
            pv = x0 @ pq.Px.T
            pre_ok = np.all((pv <= pq.pre_hi[None, :]) & (pv >= pq.pre_lo[None, :]), axis=1)
        qnorm = (np.abs(q) * self.Dinv[None, :]).max(1) / f(self.c)

        rho, alpha = f(self.rho), f(self.alpha)
        wg = np.zeros((Bn, m), dtype=f)
        wb = np.zeros((Bn, n), dtype=f)
        cg = np.clip(wg, lo, hi)
        cb = np.clip(wb, lbs, ubs)
        q64 = ((x0 - xref[None, :]) @ self.Fs.T)
        GsT = self.Gs.astype(np.float64) if self.precise else self.Gs
        KinvT = self.Kinv64 if self.precise else self.Kinv
        lamT = self.lam.astype(np.float64) if self.precise else self.lam
        tt = np.float64 if self.precise else f

        def make_t(cg, wg, cb, wb):
            return -(q64 if self.precise else q) + tt(self.rho) * ((2 * cg - wg).astype(tt) @ GsT + (2 * cb - wb).astype(tt) * lamT)

        t = make_t(cg, wg, cb, wb)
        status = np.full(Bn, 2, dtype=np.int32)
        status[~pre_ok] = 1
        iters = np.zeros(Bn, dtype=np.int32)
        xout = np.zeros((Bn, n), dtype=f)
        self.w_out = np.zeros((Bn, m + n), dtype=f)        # final w (scaled), kept for polish experiments
        live = pre_ok.copy()
        for it in range(1, max_iter + 1):
            xt = (t @ KinvT).astype(f)                                      # Kinv symmetric
            zg = xt @ self.Gs.T
            zb = xt * self.lam
            cg0 = np.clip(wg, lo, hi)
            cb0 = np.clip(wb, lbs, ubs)
            wg = wg + alpha * (zg - cg0)
            wb = wb + alpha * (zb - cb0)
            cg = np.clip(wg, lo, hi)
            cb = np.clip(wb, lbs, ubs)
            t = make_t(cg, wg, cb, wb)
            if it % check_every == 0 or it == max_iter:
                # residuals, unscaled
                rp = np.maximum((np.abs(zg - cg) * self.Einv_g).max(1) if m else 0, (np.abs(zb - cb) * self.Einv_b).max(1))
                nz = np.maximum.reduce([(np.abs(zg) * self.Einv_g).max(1) if m else np.zeros(Bn, f),
                                        (np.abs(cg) * self.Einv_g).max(1) if m else np.zeros(Bn, f),
                                        (np.abs(zb) * self.Einv_b).max(1), (np.abs(cb) * self.Einv_b).max(1)])
                dg = (2 - alpha) * cg0 - (1 - alpha) * zg - cg
                db = (2 - alpha) * cb0 - (1 - alpha) * zb - cb
                t2 = rho * (dg @ self.Gs + db * self.lam)
                rd = (np.abs(t2) * self.Dinv).max(1) / f(self.c)
                eg = alpha * zg + (1 - alpha) * cg0 - cg
                eb = alpha * zb + (1 - alpha) * cb0 - cb
                eg = np.where(np.isfinite(lo), eg, np.maximum(eg, 0))
                eb = np.where(np.isfinite(lbs)[None, :], eb, np.maximum(eb, 0))
                eb = np.where(np.isfinite(ubs)[None, :], eb, np.minimum(eb, 0))
                t3 = eg @ self.Gs + eb * self.lam
                ndy = np.maximum((np.abs(eg) / self.Einv_g).max(1) if m else 0, (np.abs(eb) / self.Einv_b).max(1))
                with np.errstate(invalid='ignore'):
                    sup = (np.where(eg > 0, hi * eg, np.where(eg < 0, np.where(np.isfinite(lo), lo, 0) * eg, 0)).sum(1)
                           + np.where(eb > 0, np.where(np.isfinite(ubs), ubs, 0) * eb,
                                      np.where(np.isfinite(lbs), lbs, 0) * eb).sum(1))
                cert = ((np.abs(t3) * self.Dinv).max(1) <= eps_inf * ndy) & (sup <= -eps_inf * ndy) & (ndy > 0)
                solved = (rp <= eps_abs + eps_rel * nz) & (rd <= eps_abs + eps_rel * qnorm)
                newly = live & (solved | cert)
                status[newly & solved] = 0
                status[newly & ~solved] = 1
                iters[newly] = it
                xout[newly] = xt[newly]
                self.w_out[newly] = np.hstack((wg, wb))[newly]
                live &= ~newly
                if verbose and it % (check_every * 10) == 0:
                    print(it, live.sum())
                if not live.any():
                    break
        xout[live] = xt[live]
        self.w_out[live] = np.hstack((wg, wb))[live]
        self.lo_out, self.hi_out = np.hstack((lo, np.broadcast_to(lbs, (Bn, n)))), np.hstack((hi, np.broadcast_to(ubs, (Bn, n))))
        iters[live] = max_iter
        u = xout.astype(float) * self.D[None, :]
        return u, status, iters
