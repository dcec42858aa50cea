"""The multiplier map an anchor of a seeded solve exports (csrc/qp_polish.cu, "explicit-MPC form"), restated in numpy: on
the anchor's critical region the multipliers of the active rows are affine in the state, lambda(x0) = Lam [x0; 1], with
Lam obtained from five right-hand sides through the Schur complement of the anchor's active set.  A follower that lies in
the same region gets its exact solution from Lam alone; one that does not fails the KKT certificate."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_env, make_controller


@pytest.fixture(scope="module")
def setup():
    from carmpc_b200.batch import BatchQP
    from oracle import carmpc_oracle as orc
    c = make_controller(make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0]), 20)
    pq = BatchQP.from_controller(c).pq
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    return pq, orc.CondensedQP("RoadOneCarEnv", 20, Ab), np.array(c.goal, dtype=float)


def _tables(pq):
    n = pq.n
    A = np.vstack((pq.G, np.eye(n)))
    Hinv = np.linalg.inv(pq.H)
    AH = A @ Hinv
    return A, Hinv, AH, AH @ A.T, -Hinv @ pq.F                      # ..., Uu: u_unc = Uu (x0 - xref)


def _kkt(pq, tables, x0, xref, act, sgn, lam):
    """u and the certificate for multipliers lam on the rows act (sign sgn): (u, certified)."""
    A, Hinv, AH, AHA, Uu = tables
    u = Uu @ (x0 - xref) - AH[act].T @ lam
    Au = A @ u
    hi = np.hstack((pq.hi - pq.Gx @ x0, pq.ub))
    lo = np.hstack((pq.lo - pq.Gx @ x0, pq.lb))
    viol = np.maximum(Au - hi, lo - Au)
    viol[~np.isfinite(viol)] = -np.inf
    signs_ok = np.all(lam * sgn >= -1e-9 * (1 + np.abs(lam).max(initial=0.0)))
    return u, bool(viol.max() <= 1e-8 and signs_ok)


def _active_set(pq, oq, x0, xref):
    from oracle import carmpc_oracle as orc
    ue, obj, st, polished, slack = orc.qp_solve_exact(oq, x0[None, :], xref)
    assert st[0] == 0 and polished[0]
    u = ue[0]
    A = np.vstack((pq.G, np.eye(pq.n)))
    Au = A @ u
    hi = np.hstack((pq.hi - pq.Gx @ x0, pq.ub))
    lo = np.hstack((pq.lo - pq.Gx @ x0, pq.lb))
    up, dn = np.abs(Au - hi) <= 1e-7, np.abs(Au - lo) <= 1e-7
    act = np.flatnonzero(up | dn)
    return u, obj[0], act, np.where(up[act], 1.0, -1.0)


def test_multiplier_map_reproduces_the_exact_solution_inside_the_region(setup):
    pq, oq, xref = setup
    tables = _tables(pq)
    A, Hinv, AH, AHA, Uu = tables
    anchor = np.array([14.0, 0.6, 0.05, 2.5])
    u_a, obj_a, act, sgn = _active_set(pq, oq, anchor, xref)
    assert 3 <= len(act) <= 32
    # rhs_a(x0) = (A u_unc)_a + shift_a(x0) - bound_a  is affine in x0:  columns = d/dx0, last column = constant
    AUu = A @ Uu
    Gx_full = np.vstack((pq.Gx, np.zeros((pq.n, 4))))
    bound = np.where(sgn > 0, np.hstack((pq.hi, pq.ub))[act], np.hstack((pq.lo, pq.lb))[act])
    rhs_cols = np.column_stack((AUu[act] + Gx_full[act], -(AUu[act] @ xref) - bound))
    M = AHA[np.ix_(act, act)]
    Lam = np.linalg.lstsq(M, rhs_cols, rcond=1e-11)[0]                     # |act| x 5
    # the anchor itself
    lam = Lam @ np.append(anchor, 1.0)
    u, ok = _kkt(pq, tables, anchor, xref, act, sgn, lam)
    assert ok and np.abs(u - u_a).max() <= 1e-8
    # followers: small steps stay in the critical region and are certified by the map alone; the solution is the exact one
    inside = 0
    for step in ([0.0, 0.02, 0, 0], [0.05, 0, 0, 0], [0, -0.03, 0.002, 0], [0.1, 0.05, 0, 0.02], [3.0, 1.2, 0.1, -1.5]):
        x0 = anchor + np.array(step)
        lam = Lam @ np.append(x0, 1.0)
        u, ok = _kkt(pq, tables, x0, xref, act, sgn, lam)
        u_e, obj_e, act_e, _ = _active_set(pq, oq, x0, xref)
        same_region = set(act_e) == set(act)
        if ok:
            inside += 1
            assert np.abs(u - u_e).max() <= 1e-7, "a certified follower must carry the exact optimum"
            q = pq.F @ (x0 - xref)
            assert abs(0.5 * u @ pq.H @ u + q @ u - obj_e) <= 1e-7 * max(1.0, abs(obj_e))
        else:
            assert not same_region, "a follower of the anchor's own critical region must be certified by the map"
    assert inside >= 2
