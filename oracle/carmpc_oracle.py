"""CPU oracle: a plain numpy restatement of CarMPC's batch-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``carmpc_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and
there only as the checker or as the CPU baseline being timed.

Pinning status
--------------
* Model, prediction / cost matrices, constraint stacks, simulator, observer, LQR step: PINNED against
  golden vectors produced by importing the unmodified reference (``tests/golden/gen_golden.py``).
* Terminal-set membership: PINNED against the reference's shipped ``terminal_sets/*.npy`` and its own
  grid expression (434 members, per-v counts [49, 70, 84, 91, 91, 49], 49 exact ties).
* QP solutions / feasibility flags: **parity unpinned**.  The reference solves its QPs with
  ``cvxpy.Problem.solve()`` (``lib/mpc.py:334-335, 477-478``), i.e. OSQP through cvxpy; neither package is in
  the reference tree nor installed here (no requirements file pins a version), and the reference ships
  no QP test or golden result.  The oracle therefore solves the *same optimisation problem* exactly
  (float64 ADMM to 1e-10, then an active-set KKT solve whose optimality conditions are verified; exact
  LP phase-1 for the feasibility flag) and the tolerances of BASELINE.json apply against that.

Every function cites the reference lines it follows.  Pure numpy / scipy; loops where clarity wins.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_discrete_are
from scipy.optimize import linprog

# ----------------------------------------------------------------------------------------------------
# constants  (lib/configuration.py:1-21, lib/simulator.py:5-13, lib/mpc.py:84-93, 387-404)
# ----------------------------------------------------------------------------------------------------
DT = 0.2
LIN_STATE = np.array([0.0, 0.0, 0.0, 3.0])
LIN_INPUT = np.array([0.0, 0.0])
Q = np.diag([5.0, 5.0, 10.0, 10.0])
R = np.diag([10.0, 100.0])
L1 = 3.5
U_UPPER = np.array([2.0, np.pi / 8])
U_LOWER = -U_UPPER
BUILTIN_STATE_ROWS = [([0, 0, 1, 0], np.pi / 8), ([0, 0, -1, 0], np.pi / 8), ([0, 0, 0, 1], 5.0), ([0, 0, 0, -1], 1.0)]
C_OUT = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 0, 1.0]])
L_OBS = np.array([[3.00000000e-01, 2.00284502e-16, 2.00000000e-01],
                  [3.48682242e+00, 1.90000000e+00, 3.86733776e-01],
                  [1.74341121e+00, 8.00000000e-01, 1.93366888e-01],
                  [-3.36822969e-16, -2.07305381e-16, 3.00000000e-01]])

ENV_ROWS = {   # lib/environments.py:56-57, 86-87, 123-124 ; default goals :63, :91, :128
    "RoadEnv": ([([0, 1, 0, 0], 3.0), ([0, -1, 0, 0], 3.0)], [30, 1.5, 0, 0]),
    "RoadOneCarEnv": ([([0, 1, 0, 0], 3.0), ([0, -1, 0, 0], 3.0), ([1, 0, 0, 0], 30.0)], [29.9, -1.5, 0, 0]),
    "RoadMultipleCarsEnv": ([([0, 1, 0, 0], 3.0), ([0, -1, 0, 0], 3.0), ([-0.25, 1, 0, 0], -2.0),
                             ([0.25, -1, 0, 0], 6.25)], [30, 1.5, 0, 0]),
}


# ----------------------------------------------------------------------------------------------------
# model  (lib/mpc.py:127-180, :70, :81)
# ----------------------------------------------------------------------------------------------------
def bicycle_model(lin_state=LIN_STATE, lin_input=LIN_INPUT, dt=DT):
    """Forward-Euler discretisation of the Jacobians of the kinematic bicycle."""
    _, _, psi, v = lin_state
    _, delta = lin_input
    Ac = np.zeros((4, 4))
    Ac[0, 2] = -v * np.sin(psi)
    Ac[0, 3] = np.cos(psi)
    Ac[1, 2] = v * np.cos(psi)
    Ac[1, 3] = np.sin(psi)
    Ac[2, 3] = np.tan(delta) / L1
    Bc = np.zeros((4, 2))
    Bc[2, 1] = v / L1 * 1 / np.cos(delta) ** 2
    Bc[3, 0] = 1.0
    return dt * Ac + np.eye(4), dt * Bc


def lqr(A, B):
    P = solve_discrete_are(A, B, Q, R)
    K = -np.linalg.inv(R + B.T @ P @ B) @ B.T @ P @ A
    return P, K


# ----------------------------------------------------------------------------------------------------
# prediction and cost matrices  (lib/matrix_gen.py:6-72)
# ----------------------------------------------------------------------------------------------------
def predmod(A, B, N):
    nx, nu = B.shape
    T = np.zeros(((N + 1) * nx, nx))
    S = np.zeros(((N + 1) * nx, N * nu))
    for i in range(N + 1):
        T[i * nx:(i + 1) * nx] = np.linalg.matrix_power(A, i)
        for j in range(N):
            if i - j - 1 >= 0:
                S[i * nx:(i + 1) * nx, j * nu:(j + 1) * nu] = np.linalg.matrix_power(A, i - j - 1) @ B
    return T, S


def costgen(P, T, S, N):
    nx, nu = 4, 2
    Qh = np.zeros(((N + 1) * nx, (N + 1) * nx))
    Rh = np.zeros((N * nu, N * nu))
    for k in range(N + 1):
        Qh[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] = Q if k < N else P
    for k in range(N):
        Rh[k * nu:(k + 1) * nu, k * nu:(k + 1) * nu] = R
    return Rh + S.T @ Qh @ S, S.T @ Qh @ T


# ----------------------------------------------------------------------------------------------------
# constraint stacks and the condensed QP  (lib/mpc.py:196-253, :318-332 )
# ----------------------------------------------------------------------------------------------------
def state_rows(env_name):
    return BUILTIN_STATE_ROWS + ENV_ROWS[env_name][0]


def state_constraint(env_name, N):
    rows = state_rows(env_name)
    A = np.zeros((len(rows) * N, (N + 1) * 4))
    b = np.zeros(len(rows) * N)
    for r, (a, bb) in enumerate(rows):
        for i in range(N):
            A[r * N + i, (i + 1) * 4:(i + 2) * 4] = a
            b[r * N + i] = bb
    return A, b


def input_constraint(N):
    n = 2 * N
    return np.vstack((np.eye(n), -np.eye(n))), np.hstack((np.tile(U_UPPER, N), np.tile(-U_LOWER, N)))


def terminal_constraint(term_Ab, N):
    A = np.zeros((len(term_Ab), (N + 1) * 4))
    A[:, -4:] = term_Ab[:, :4]
    return A, term_Ab[:, 4].copy()


class CondensedQP:
    """min 1/2 u'Hu + (h (x0 - xref))'u   s.t.  G u <= w - Gx x0   (all rows one-sided, as the reference states them)."""

    def __init__(self, env_name, N, term_Ab, use_terminal=True, use_input=True, use_state=True):
        self.N, self.n = N, 2 * N
        self.A, self.B = bicycle_model()
        self.P, self.K = lqr(self.A, self.B)
        self.T, self.S = predmod(self.A, self.B, N)
        self.H, self.h = costgen(self.P, self.T, self.S, N)
        blocks_G, blocks_Gx, blocks_w = [], [], []
        if use_terminal:
            At, bt = terminal_constraint(term_Ab, N)
            blocks_G.append(At @ self.S)
            blocks_Gx.append(At @ self.T)
            blocks_w.append(bt)
        if use_input:
            Ai, bi = input_constraint(N)
            blocks_G.append(Ai)
            blocks_Gx.append(np.zeros((len(bi), 4)))
            blocks_w.append(bi)
        if use_state:
            As, bs = state_constraint(env_name, N)
            blocks_G.append(As @ self.S)
            blocks_Gx.append(As @ self.T)
            blocks_w.append(bs)
        self.G = np.vstack(blocks_G) if blocks_G else np.zeros((0, self.n))
        self.Gx = np.vstack(blocks_Gx) if blocks_Gx else np.zeros((0, 4))
        self.w = np.hstack(blocks_w) if blocks_w else np.zeros(0)

    def rhs(self, x0):
        """Upper bounds  w - Gx x0  for a batch x0 (B, 4)."""
        return self.w[None, :] - np.atleast_2d(x0) @ self.Gx.T

    def lin(self, x0, xref):
        return (np.atleast_2d(x0) - np.asarray(xref)[None, :]) @ self.h.T

    def objective(self, u, x0, xref):
        q = self.lin(x0, xref)
        return 0.5 * np.einsum('bi,ij,bj->b', u, self.H, u) + np.einsum('bi,bi->b', q, u)


# ----------------------------------------------------------------------------------------------------
# exact feasibility  (sampled equivalent of lib/in_adm_set.py ; cvxpy reports +-inf -> lib/mpc.py:336)
# ----------------------------------------------------------------------------------------------------
def qp_feasible_lp(qp: CondensedQP, x0, margin_out=None):
    """Exact flag per sample: max t s.t. G u + t <= w - Gx x0.  Feasible iff t* >= 0.  Returns (flag, t*)."""
    ub = qp.rhs(x0)
    m, n = qp.G.shape
    A_ub = np.hstack((qp.G, np.ones((m, 1))))
    c = np.zeros(n + 1)
    c[-1] = -1.0
    flags = np.zeros(len(ub), dtype=bool)
    slack = np.zeros(len(ub))
    for i, b in enumerate(ub):
        res = linprog(c, A_ub=A_ub, b_ub=b, bounds=[(None, None)] * n + [(None, 1.0)], method="highs")
        slack[i] = -res.fun if res.status == 0 else -np.inf
        flags[i] = res.status == 0 and slack[i] >= 0.0
    return flags, slack


# ----------------------------------------------------------------------------------------------------
# exact QP solution: float64 ADMM (OSQP iteration) + verified active-set polish
# ----------------------------------------------------------------------------------------------------
def qp_solve_admm(qp: CondensedQP, x0, xref, rho=30.0, sigma=1e-6, alpha=1.6, eps=1e-10, max_iter=20000,
                  check_every=25, eps_inf=1e-7, return_iters=False):
    """Batched OSQP-style ADMM in float64 on  min 1/2 u'Hu + q'u, G u <= ub.

    Returns u (B, n), y (B, m), status (B,) with 0 solved, 1 primal infeasible (certificate), 2 max_iter.
    """
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    Bn = len(x0)
    G, H = qp.G, qp.H
    m, n = G.shape
    ub = qp.rhs(x0)
    q = qp.lin(x0, xref)
    Kinv = np.linalg.inv(H + sigma * np.eye(n) + rho * G.T @ G)
    x = np.zeros((Bn, n))
    z = np.zeros((Bn, m))
    y = np.zeros((Bn, m))
    status = np.full(Bn, 2, dtype=np.int32)
    iters = np.zeros(Bn, dtype=np.int32)
    active = np.arange(Bn)
    out_x = np.zeros((Bn, n))
    out_y = np.zeros((Bn, m))
    for it in range(1, max_iter + 1):
        rhs = sigma * x - q[active] + (rho * z - y) @ G
        xt = rhs @ Kinv
        zt = xt @ G.T
        x_new = alpha * xt + (1 - alpha) * x
        zh = alpha * zt + (1 - alpha) * z
        z_new = np.minimum(zh + y / rho, ub[active])
        y_new = y + rho * (zh - z_new)
        dy = y_new - y
        x, z, y = x_new, z_new, y_new
        if it % check_every == 0 or it == max_iter:
            Gx_ = x @ G.T
            r_prim = np.abs(Gx_ - z).max(1)
            Hx = x @ H
            Gty = y @ G
            r_dual = np.abs(Hx + q[active] + Gty).max(1)
            e_prim = eps + eps * np.maximum(np.abs(Gx_).max(1), np.abs(z).max(1))
            e_dual = eps + eps * np.maximum.reduce([np.abs(Hx).max(1), np.abs(Gty).max(1), np.abs(q[active]).max(1)])
            solved = (r_prim <= e_prim) & (r_dual <= e_dual)
            ndy = np.abs(dy).max(1)
            cert = (np.abs(dy @ G).max(1) <= eps_inf * ndy) & \
                   ((ub[active] * np.maximum(dy, 0)).sum(1) <= -eps_inf * ndy) & (ndy > 0)
            done = solved | cert
            if done.any():
                idx = active[done]
                status[idx] = np.where(solved[done], 0, 1)
                iters[idx] = it
                out_x[idx] = x[done]
                out_y[idx] = y[done]
                keep = ~done
                active, x, z, y = active[keep], x[keep], z[keep], y[keep]
                if len(active) == 0:
                    break
    if len(active):
        out_x[active], out_y[active] = x, y
        iters[active] = max_iter
    if return_iters:
        return out_x, out_y, status, iters
    return out_x, out_y, status


def qp_polish(qp: CondensedQP, x0, xref, u, y, tol=1e-9, max_rounds=30):
    """Active-set refinement of one ADMM solution: solve the equality-constrained KKT system on the guessed
    active set, repair the set (drop negative multipliers, add the most violated row) until the KKT
    conditions hold to ``tol``.  Returns (u, lambda, ok)."""
    G, H = qp.G, qp.H
    ub = qp.rhs(x0)[0]
    q = qp.lin(x0, xref)[0]
    n = qp.n
    act = list(np.flatnonzero((y > 1e-7) | (G @ u - ub > -1e-7)))
    lam = np.zeros(len(ub))
    for _ in range(max_rounds):
        # keep a linearly independent subset (rank-revealing greedy)
        sel = []
        for r in act:
            cand = sel + [r]
            if np.linalg.matrix_rank(G[cand], tol=1e-10) == len(cand):
                sel = cand
        k = len(sel)
        KKT = np.zeros((n + k, n + k))
        KKT[:n, :n] = H
        KKT[:n, n:] = G[sel].T
        KKT[n:, :n] = G[sel]
        sol = np.linalg.solve(KKT, np.hstack((-q, ub[sel])))
        u_new, mult = sol[:n], sol[n:]
        lam = np.zeros(len(ub))
        lam[sel] = mult
        viol = G @ u_new - ub
        worst = int(np.argmax(viol))
        neg = [r for r, mu in zip(sel, mult) if mu < -tol]
        if viol[worst] <= tol and not neg:
            return u_new, lam, True
        act = [r for r in sel if r not in neg]
        if viol[worst] > tol and worst not in act:
            act.append(worst)
    return u, lam, False


def qp_solve_exact(qp: CondensedQP, x0, xref=None, rho=30.0):
    """Reference answer for a batch: u (B, n), objective (B,), status (B,) (0 solved / 1 infeasible), polished (B,).

    Feasibility comes from the exact LP; solutions from ADMM + verified polish.
    """
    x0 = np.atleast_2d(np.asarray(x0, dtype=float))
    xref = np.zeros(4) if xref is None else np.asarray(xref, dtype=float)
    feas, slack = qp_feasible_lp(qp, x0)
    u = np.full((len(x0), qp.n), np.nan)
    obj = np.full(len(x0), np.inf)
    polished = np.zeros(len(x0), dtype=bool)
    idx = np.flatnonzero(feas)
    if len(idx):
        ua, ya, st = qp_solve_admm(qp, x0[idx], xref, rho=rho)
        for k, i in enumerate(idx):
            up, _, ok = qp_polish(qp, x0[i:i + 1], xref, ua[k], ya[k])
            u[i] = up if ok else ua[k]
            polished[i] = ok
        obj[idx] = qp.objective(u[idx], x0[idx], xref)
    return u, obj, np.where(feas, 0, 1).astype(np.int32), polished, slack


# ----------------------------------------------------------------------------------------------------
# terminal-set membership  (lib/terminal_set.py:107-113) and its rollout form (:53-59, :198-200)
# ----------------------------------------------------------------------------------------------------
def membership_pointwise(Ab, points):
    """The reference expression, one point at a time: np.all(A @ point <= b)."""
    A, b = Ab[:, :4], Ab[:, 4]
    return np.array([bool(np.all(A @ p <= b)) for p in points])


def membership(Ab, x, y, psi, v, chunk=1 << 20):
    """Vectorised form of the same test on SoA grids; returns (member, min margin)."""
    A, b = Ab[:, :4], Ab[:, 4]
    n = len(x)
    out = np.zeros(n, dtype=bool)
    margin = np.zeros(n)
    for s in range(0, n, chunk):
        P = np.stack((x[s:s + chunk], y[s:s + chunk], psi[s:s + chunk], v[s:s + chunk]), axis=1)
        r = P @ A.T
        out[s:s + chunk] = (r <= b).all(1)
        margin[s:s + chunk] = (b - r).min(1)
    return out, margin


def normalise_rows(A, b):
    nrm = np.sqrt((A * A).sum(1))
    return A / nrm[:, None], b / nrm


def rollout_setup(env_name, goal):
    """A_k, goal-shifted unit-norm state rows and t = 0 input rows of calc_terminal_set (:145-161, :198-200)."""
    A, B = bicycle_model()
    _, K = lqr(A, B)
    rows = state_rows(env_name)
    Ac = np.array([r for r, _ in rows], dtype=float)
    bc = np.array([bb for _, bb in rows], dtype=float)
    Ac, bc = normalise_rows(Ac, bc)
    goal = np.asarray(goal, dtype=float)
    bc = bc - Ac @ goal                                   # translation(-goal)
    Ai = np.vstack((np.eye(2), -np.eye(2))) @ K
    bi = np.hstack((U_UPPER, -U_LOWER))
    Ai, bi = normalise_rows(Ai, bi)
    return A + B @ K, K, Ac, bc, Ai, bi


def rollout_membership(env_name, goal, k_steps, x, y, psi, v, input_every_step=False):
    """Sampled form of the terminal set: e_0 = p - goal, e_{t+1} = A_k e_t; state rows checked for t = 0..k_steps,
    input rows at t = 0 (reference semantics) or at every step.  Returns (member, first violated step or -1, margin)."""
    Ak, K, Ac, bc, Ai, bi = rollout_setup(env_name, goal)
    e = np.stack((x, y, psi, v), axis=1) - np.asarray(goal, dtype=float)[None, :]
    n = len(e)
    member = np.ones(n, dtype=bool)
    first = np.full(n, -1, dtype=np.int32)
    margin = np.full(n, np.inf)
    for t in range(k_steps + 1):
        r = bc[None, :] - e @ Ac.T
        ok = (r >= 0).all(1)
        margin = np.minimum(margin, r.min(1))
        if t == 0 or input_every_step:
            ri = bi[None, :] - e @ Ai.T
            ok &= (ri >= 0).all(1)
            margin = np.minimum(margin, ri.min(1))
        newly = member & ~ok
        first[newly] = t
        member &= ok
        e = e @ Ak.T
    return member, first, margin


# ----------------------------------------------------------------------------------------------------
# plant, observer and the closed loops  (lib/simulator.py:51-69 ; lib/mpc.py:448 ; examples/run_MPC*.py)
# ----------------------------------------------------------------------------------------------------
def plant_step(state, u, dt=DT):
    """Batched forward-Euler bicycle: state (B, 4), u (B, 2)."""
    psi, v = state[:, 2], state[:, 3]
    rate = np.stack((v * np.cos(psi), v * np.sin(psi), v / L1 * np.tan(u[:, 1]), u[:, 0]), axis=1)
    return rate * dt + state


def observer_step(A, B, xhat, u_prev, y):
    return xhat @ A.T + u_prev @ B.T + (y - xhat @ C_OUT.T) @ L_OBS.T


def closed_loop(qp: CondensedQP, x_init, goal, steps, output_feedback, solve):
    """Monte-Carlo closed loop in the order of examples/run_MPCOutputFB.py:29-41 / run_MPCStateFB.py:29-39:
    plant step with the previous input (first [0, 0]), then controller step.  ``solve(x0) -> (u0, status)``.
    A run stops being updated at its first infeasible step (the reference raises).  Returns final states,
    fail_step (-1 = none) and the state trajectory."""
    x = np.array(x_init, dtype=float)
    Bn = len(x)
    xhat = x.copy()
    u = np.zeros((Bn, 2))
    fail = np.full(Bn, -1, dtype=np.int32)
    traj = np.zeros((steps, Bn, 4))
    A, Bm = qp.A, qp.B
    for k in range(steps):
        alive = fail < 0
        x[alive] = plant_step(x[alive], u[alive])
        if output_feedback:
            xhat[alive] = observer_step(A, Bm, xhat[alive], u[alive], x[alive] @ C_OUT.T)
            est = xhat
        else:
            est = x
        u_new, st = solve(est[alive])
        idx = np.flatnonzero(alive)
        bad = st != 0
        fail[idx[bad]] = k
        u[idx[~bad]] = u_new[~bad]
        traj[k] = x
    return x, fail, traj
