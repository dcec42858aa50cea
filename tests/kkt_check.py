"""Solver-independent check of QP results against the REFERENCE's problem definition.

Test infrastructure.  The matrices come from ``tests/golden/qp_<env>_N<k>.npz``, which ``gen_golden.py`` wrote from
the unmodified reference controller object (H, h, T, S and the three constraint stacks exactly as
``MPCStateFB.step`` uses them, lib/mpc.py:318-332).  Nothing here imports ``oracle/`` or any QP solver: a result is
accepted because it satisfies the optimality conditions of the reference's QP in float64,

    x = T x0 + S u ;  A_term x <= b_term ;  A_in u <= b_in ;  A_state x <= b_state                 (primal, :324-332)
    H u + h (x0 - goal) + G' lambda = 0 ,  lambda >= 0 ,  lambda_i = 0 on inactive rows             (stationarity)

(the problem is strictly convex, H > 0, so a KKT point is THE optimum), and an "infeasible" flag is accepted because a
phase-1 LP (scipy / HiGHS) over the same rows has no solution.
"""
import os

import numpy as np
from scipy.optimize import linprog, nnls

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ReferenceQP:
    def __init__(self, env_name: str, N: int):
        d = np.load(os.path.join(GOLDEN, f"qp_{env_name}_N{N}.npz"))
        self.N, self.n = N, 2 * N
        self.H, self.h, self.T, self.S = d["H"], d["h"], d["T"], d["S"]
        self.P, self.Q, self.R = d["P"], d["Q"], d["R"]
        self.goal = d["goal"]
        At, Ai, As = d["At"], d["Ai"], d["As"]
        # the three `constraints += [...]` of lib/mpc.py:324-332, with x = T x0 + S u substituted (:319)
        self.G = np.vstack((At @ self.S, Ai, As @ self.S))
        self.Gx = np.vstack((At @ self.T, np.zeros((len(Ai), 4)), As @ self.T))
        self.w = np.hstack((d["bt"], d["bi"], d["bs"]))
        self.row_scale = np.maximum(1.0, np.abs(self.G).sum(1))

    def rhs(self, x0):
        return self.w[None, :] - np.atleast_2d(x0) @ self.Gx.T

    def gradient(self, u, x0, xref=None):
        xref = self.goal if xref is None else np.asarray(xref, dtype=float)
        return np.atleast_2d(u) @ self.H + (np.atleast_2d(x0) - xref) @ self.h.T       # H symmetric

    def objective(self, u, x0, xref=None):
        """The reference's objective value (lib/mpc.py:321): 1/2 u'Hu + (h (x0 - goal))'u."""
        xref = self.goal if xref is None else np.asarray(xref, dtype=float)
        u = np.atleast_2d(u)
        return 0.5 * np.einsum("bi,ij,bj->b", u, self.H, u) + np.einsum("bi,bi->b", (np.atleast_2d(x0) - xref) @ self.h.T, u)

    def kkt(self, u, x0, xref=None, act_tol=1e-7):
        """Per sample: (primal residual, stationarity residual relative to max(1, |gradient|_inf), number of active
        rows).  Multipliers are the non-negative least-squares fit on the active rows, so a small stationarity residual
        proves that non-negative multipliers exist."""
        u, x0 = np.atleast_2d(u), np.atleast_2d(x0)
        viol = u @ self.G.T - self.rhs(x0)
        primal = np.maximum(viol, 0.0).max(1)
        grad = self.gradient(u, x0, xref)
        stat = np.zeros(len(u))
        nact = np.zeros(len(u), dtype=int)
        for i in range(len(u)):
            act = np.flatnonzero(viol[i] >= -act_tol * self.row_scale)
            nact[i] = len(act)
            scale = max(1.0, np.abs(grad[i]).max())
            if len(act) == 0:
                stat[i] = np.abs(grad[i]).max() / scale
                continue
            lam, _ = nnls(self.G[act].T, -grad[i], maxiter=50 * self.n)
            stat[i] = np.abs(self.G[act].T @ lam + grad[i]).max() / scale
        return primal, stat, nact

    def feasibility_slack(self, x0):
        """Phase-1 LP per state: max s such that G u + s <= rhs (s capped at 1).  s* < 0: the reference's QP has no
        feasible point (cvxpy returns inf, lib/mpc.py:336-338 raises); |s*| <= 1e-6 is BASELINE's boundary band."""
        rhs = self.rhs(x0)
        m, n = self.G.shape
        A = np.hstack((self.G, np.ones((m, 1))))
        c = np.zeros(n + 1)
        c[-1] = -1.0
        out = np.zeros(len(rhs))
        for i, b in enumerate(rhs):
            if not np.all(np.isfinite(b)):
                out[i] = -np.inf
                continue
            res = linprog(c, A_ub=A, b_ub=b, bounds=[(None, None)] * n + [(None, 1.0)], method="highs")
            out[i] = -res.fun if res.status == 0 else -np.inf
        return out


def assert_results_satisfy_reference_qp(ref: ReferenceQP, x0, u_full, status, objective=None, xref=None, n_lp=400,
                                        seed=0, primal_tol=1e-8, stat_tol=1e-6):
    """Every solved sample is a KKT point of the reference's QP; a sample of the infeasible flags is confirmed by the
    phase-1 LP (outside the 1e-6 band); returns a summary dict for the test log."""
    x0, status = np.atleast_2d(x0), np.asarray(status)
    assert not (status == 2).any(), f"{int((status == 2).sum())} samples undecided (max_iter)"
    ok = status == 0
    primal, stat, nact = ref.kkt(u_full[ok], x0[ok], xref)
    worst_p, worst_s = (float(primal.max()), float(stat.max())) if ok.any() else (0.0, 0.0)
    assert worst_p <= primal_tol, f"primal residual {worst_p:.3e} against the reference rows"
    assert worst_s <= stat_tol, f"stationarity residual {worst_s:.3e} (sample {int(np.argmax(stat))}, {int(nact[np.argmax(stat)])} active rows)"
    if objective is not None and ok.any():
        want = ref.objective(u_full[ok], x0[ok], xref)
        rel = np.abs(np.asarray(objective)[ok] - want) / np.maximum(1.0, np.abs(want))
        assert rel.max() <= 1e-9, f"objective differs from the reference expression by {rel.max():.3e}"
    bad = np.flatnonzero(~ok)
    rng = np.random.default_rng(seed)
    pick = bad if len(bad) <= n_lp else rng.choice(bad, size=n_lp, replace=False)
    slack = ref.feasibility_slack(x0[pick])
    wrong = pick[slack > 1e-6]
    assert len(wrong) == 0, f"{len(wrong)} states flagged infeasible have a strictly feasible point (slack up to {slack.max():.3e})"
    return {"solved": int(ok.sum()), "infeasible": int(len(bad)), "lp_checked": int(len(pick)),
            "in_band": int((np.abs(slack) <= 1e-6).sum()), "primal": worst_p, "stationarity": worst_s,
            "max_active": int(nact.max()) if ok.any() else 0}
