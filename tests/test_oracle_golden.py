"""The CPU oracle against golden vectors produced by the unmodified reference (tests/golden/gen_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, FIXTURES, K_STAR, golden
from oracle import carmpc_oracle as orc
from oracle import c_oracle


def test_model_matches_reference():
    g = golden("model.npz")
    A, B = orc.bicycle_model()
    P, K = orc.lqr(A, B)
    np.testing.assert_array_equal(A, g["A"])
    np.testing.assert_array_equal(B, g["B"])
    np.testing.assert_allclose(P, g["P"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(K, g["K"], rtol=0, atol=1e-11)
    np.testing.assert_array_equal(orc.L_OBS, g["L"])
    np.testing.assert_array_equal(orc.C_OUT, g["C"])
    np.testing.assert_array_equal(orc.U_UPPER, g["input_upper"])


@pytest.mark.parametrize("N", [1, 5, 10, 20])
def test_prediction_and_cost_matrices(N):
    g = golden(f"predmod_N{N}.npz")
    A, B = orc.bicycle_model()
    P, _ = orc.lqr(A, B)
    T, S = orc.predmod(A, B, N)
    H, h = orc.costgen(P, T, S, N)
    np.testing.assert_allclose(T, g["T"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(S, g["S"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(H, g["H"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(h, g["h"], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("N", [40, 80])
def test_prediction_digest_long_horizons(N):
    g = golden(f"predmod_digest_N{N}.npz")
    A, B = orc.bicycle_model()
    P, _ = orc.lqr(A, B)
    T, S = orc.predmod(A, B, N)
    H, h = orc.costgen(P, T, S, N)
    np.testing.assert_allclose(T[-4:], g["T_last"], rtol=1e-12)
    np.testing.assert_allclose(S[-4:], g["S_last"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.diag(H), g["H_diag"], rtol=1e-10)
    np.testing.assert_allclose(H[0], g["H_row0"], rtol=1e-10, atol=1e-7)
    np.testing.assert_allclose(np.trace(H), g["H_trace"], rtol=1e-11)
    np.testing.assert_allclose(h[:2], g["h_first"], rtol=1e-10, atol=1e-7)


@pytest.mark.parametrize("tag,env,file", [
    ("RoadEnv", "RoadEnv", "RoadEnv_30_1.5_0_0.npy"),
    ("RoadOneCarEnv", "RoadOneCarEnv", "RoadOneCarEnv_29.9_1.5_0_0.npy"),
    ("RoadMultipleCarsEnv", "RoadMultipleCarsEnv", "RoadMultipleCarsEnv_30_1.5_0_0.npy")])
@pytest.mark.parametrize("N", [1, 3, 20])
def test_constraint_stacks(tag, env, file, N):
    g = golden(f"constraints_{tag}_N{N}.npz")
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", file))
    At, bt = orc.terminal_constraint(Ab, N)
    Ai, bi = orc.input_constraint(N)
    As, bs = orc.state_constraint(env, N)
    np.testing.assert_array_equal(At, g["At"])
    np.testing.assert_array_equal(bt, g["bt"])
    np.testing.assert_array_equal(Ai, g["Ai"])
    np.testing.assert_allclose(bi, g["bi"], rtol=0, atol=0)
    np.testing.assert_array_equal(As, g["As"])
    np.testing.assert_array_equal(bs, g["bs"])


def test_plant_matches_reference_simulator():
    g = golden("simulator.npz")
    x = g["x0_0"][None, :].copy()
    for u, want in zip(g["u_0"], g["states_0"]):
        x = orc.plant_step(x, u[None, :], dt=float(g["dt_0"]))
        np.testing.assert_allclose(x[0], want, rtol=0, atol=1e-13)
    np.testing.assert_allclose(g["states_0"] @ orc.C_OUT.T, g["outputs_0"], atol=1e-13)


def test_observer_matches_reference():
    g = golden("observer.npz")
    A, B = orc.bicycle_model()
    xh = g["xhat"][0][None, :]
    for y, u, want in zip(g["y"], g["u"], g["xhat"][1:]):
        xh = orc.observer_step(A, B, xh, u[None, :], y[None, :])
        np.testing.assert_allclose(xh[0], want, rtol=0, atol=1e-12)


def test_lqr_step_matches_reference():
    g = golden("lqr.npz")
    A, B = orc.bicycle_model()
    _, K = orc.lqr(A, B)
    u = np.clip((g["x"] - np.array([30, 1.5, 0, 0])) @ K.T, orc.U_LOWER, orc.U_UPPER)
    np.testing.assert_allclose(u, g["u"], rtol=0, atol=1e-10)


# ---- membership: the reference grid of lib/terminal_set.py:96-113 ---------------------------------------------
def _config1_grid():
    g = golden("grid_config1.npz")
    xx, yy = g["xx"], g["yy"]
    x = np.tile(xx.ravel(), 6)
    y = np.tile(yy.ravel(), 6)
    v = np.repeat(np.arange(6.0), xx.size)
    return g, x, y, np.zeros_like(x), v


def test_membership_config1_known_answer():
    g, x, y, psi, v = _config1_grid()
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    want = g["member"].reshape(-1)
    assert want.sum() == 434
    assert list(g["member"].reshape(6, -1).sum(1)) == [49, 70, 84, 91, 91, 49]
    assert (g["margin"][g["member"]] == 0).sum() == 49          # exact ties: '<=' semantics are exercised
    got, _ = orc.membership(Ab, x, y, psi, v)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(orc.membership_pointwise(Ab, np.stack((x, y, psi, v), 1)[:3000]), want[:3000])
    bits, cnt = c_oracle.membership_bits(Ab, x, y, psi, v)
    assert cnt == 434
    np.testing.assert_array_equal(c_oracle.unpack_bits(bits, len(x)), want)
    np.testing.assert_array_equal(c_oracle.membership_plain(Ab, np.stack((x, y, psi, v), 1)), want)


@pytest.mark.parametrize("file", sorted(FIXTURES))
def test_c_oracle_equals_numpy_oracle_outside_boundary_band(file):
    """The C oracle fixes the summation order (fma chain); numpy's may differ in the last ulp.  They must agree on
    every sample whose margin exceeds 1e-6 (BASELINE north_star); the rest are enumerated."""
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", file))
    rng = np.random.default_rng(1)
    n = 400_000
    lo = np.array([0.0, -3.5, -0.45, -1.5])
    hi = np.array([35.0, 3.5, 0.45, 5.5])
    p = rng.uniform(lo, hi, size=(n, 4))
    got_np, margin = orc.membership(Ab, *p.T)
    bits, cnt = c_oracle.membership_bits(Ab, *p.T, threads=2)
    got_c = c_oracle.unpack_bits(bits, n)
    assert cnt == got_c.sum()
    band = np.abs(margin) <= 1e-6
    np.testing.assert_array_equal(got_c[~band], got_np[~band])
    assert band.sum() < 20


@pytest.mark.parametrize("file", sorted(FIXTURES))
def test_rollout_form_equals_hrep(file):
    """SURVEY 0.1 [probe]: state rows along the LQR rollout for t = 0..k*, inputs at t = 0 == the shipped H-rep."""
    env, goal = FIXTURES[file]
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", file))
    rng = np.random.default_rng(2)
    n = 300_000
    g = np.array(goal, dtype=float)
    p = g + rng.uniform(-1, 1, size=(n, 4)) * np.array([12.0, 2.5, 0.45, 3.0])
    member_h, margin_h = orc.membership(Ab, *p.T)
    member_r, first, margin_r = orc.rollout_membership(env, goal, K_STAR[env], *p.T)
    band = (np.abs(margin_h) <= 1e-6) | (np.abs(margin_r) <= 1e-6)
    np.testing.assert_array_equal(member_h[~band], member_r[~band])
    assert member_h.sum() > 1000
    Ak, K, Ac, bc, Ai, bi = orc.rollout_setup(env, goal)
    bits, first_c, cnt = c_oracle.rollout_bits(Ak, Ac, bc, Ai, bi, g, K_STAR[env], 0, *p.T)
    got_c = c_oracle.unpack_bits(bits, n)
    np.testing.assert_array_equal(got_c[~band], member_r[~band])
    np.testing.assert_array_equal(first_c[~band], first[~band])


def test_qp_oracle_known_answer():
    """RoadOneCarEnv goal (29.9, 1.5, 0, 0), N = 20, a state with saturated acceleration: the exact
    solution satisfies the KKT conditions and the LP feasibility flag agrees with the ADMM certificate."""
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    qp = orc.CondensedQP("RoadOneCarEnv", 20, Ab)
    goal = np.array([29.9, 1.5, 0, 0])
    rng = np.random.default_rng(0)
    x0 = rng.uniform([5, -3, -np.pi / 8, -1], [30, 3, np.pi / 8, 5], size=(24, 4))
    u, obj, status, polished, slack = orc.qp_solve_exact(qp, x0, goal)
    feas = status == 0
    assert 0 < feas.sum() < len(x0)
    assert polished[feas].all()
    # KKT of the polished points
    ub = qp.rhs(x0)
    for i in np.flatnonzero(feas):
        assert (qp.G @ u[i] - ub[i]).max() <= 1e-8
    # the ADMM certificate gives the same flags as the LP
    _, _, st = orc.qp_solve_admm(qp, x0, goal, eps=1e-7, max_iter=6000)
    np.testing.assert_array_equal(np.where(st == 0, 0, 1), status)


def test_qp_oracle_agrees_with_an_independent_solver():
    """QP parity is unpinned against the reference's cvxpy -> OSQP (not installable here); as the next best thing the
    oracle's exact solutions are cross-checked against an independent algorithm (scipy SLSQP, a sequential quadratic
    programming method) on the reference's own QP data: same objective to 1e-8 relative, same inputs to SLSQP's
    accuracy.  OSQP at its cvxpy defaults (eps 1e-5) would itself differ from both by ~1e-4."""
    from scipy.optimize import minimize
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    qp = orc.CondensedQP("RoadOneCarEnv", 20, Ab)
    goal = np.array([29.9, 1.5, 0, 0])
    rng = np.random.default_rng(0)
    x0 = rng.uniform([5, -3, -np.pi / 8, -1], [30, 3, np.pi / 8, 5], size=(10, 4))
    u, obj, status, polished, slack = orc.qp_solve_exact(qp, x0, goal)
    checked = 0
    for i in np.flatnonzero((status == 0) & polished)[:5]:
        q = qp.lin(x0[i:i + 1], goal)[0]
        ub = qp.rhs(x0[i:i + 1])[0]
        res = minimize(lambda z: 0.5 * z @ qp.H @ z + q @ z, np.zeros(qp.n), jac=lambda z: qp.H @ z + q,
                       constraints=[{"type": "ineq", "fun": lambda z: ub - qp.G @ z, "jac": lambda z: -qp.G}],
                       method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
        assert (qp.G @ res.x - ub).max() <= 1e-7
        assert abs(res.fun - obj[i]) <= 1e-8 * abs(obj[i])
        assert np.abs(res.x - u[i]).max() <= 1e-3
        assert obj[i] <= res.fun + 1e-9 * abs(obj[i])          # the oracle's point is at least as good
        checked += 1
    assert checked >= 3


def test_closed_loop_fixture_is_what_the_oracle_computes():
    """tests/golden/closed_loop_config4.npz (the oracle's 200 x 200 closed loops, generated offline because they take
    six minutes) cannot drift from the oracle: eight of its runs are re-derived live for 40 steps, both feedback modes."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_oracle_fixtures", os.path.join(GOLDEN, "gen_oracle_fixtures.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    fx = np.load(os.path.join(GOLDEN, "closed_loop_config4.npz"))
    np.testing.assert_array_equal(fx["x_init"], gen.config4_initial_states())
    pick = np.array([0, 3, 17, 42, 77, 101, 150, 199])
    stride = int(fx["stride"])
    for tag, fb in (("ofb", True), ("sfb", False)):
        x, fail, traj = gen.run_oracle_loop(fx["x_init"][pick], 40, fb)
        want_fail = fx[f"fail_{tag}"][pick]
        np.testing.assert_array_equal(fail, np.where((want_fail >= 0) & (want_fail < 40), want_fail, -1))
        np.testing.assert_allclose(traj[stride - 1::stride], fx[f"traj_{tag}"][:40 // stride, pick], rtol=0, atol=1e-9)
