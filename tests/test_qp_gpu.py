"""Parity of the batched QP path (CUDA ADMM + float64 polish, through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: first inputs within 1e-4 absolute, objectives within 1e-5 relative, feasibility flags
identical (states whose exact feasibility slack is within 1e-6 of zero are enumerated, not compared).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_env, make_controller

pytestmark = pytest.mark.gpu

U_TOL = 1e-4
OBJ_RTOL = 1e-5
FIXTURE = {"RoadOneCarEnv": "RoadOneCarEnv_29.9_1.5_0_0.npy", "RoadMultipleCarsEnv": "RoadMultipleCarsEnv_30_1.5_0_0.npy",
           "RoadEnv": "RoadEnv_30_1.5_0_0.npy"}
GOAL = {"RoadOneCarEnv": [29.9, 1.5, 0, 0], "RoadMultipleCarsEnv": None, "RoadEnv": None}


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "the GPU tests need a CUDA device"
    return torch


def _setup(env_name, N, **opts):
    from carmpc_b200.batch import BatchQP
    from oracle import carmpc_oracle as orc
    c = make_controller(make_env(env_name, GOAL[env_name]), N)
    bq = BatchQP.from_controller(c, **opts)
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", FIXTURE[env_name]))
    return c, bq, orc.CondensedQP(env_name, N, Ab)


def _states(c, n, seed, spread=(12.0, 1.4, 0.25, 2.5)):
    rng = np.random.default_rng(seed)
    g = np.array(c.goal, dtype=float)
    return g + rng.uniform(-1, 1, size=(n, 4)) * np.array(spread)


def _compare(res, x0, oq, goal, min_feasible=5):
    from oracle import carmpc_oracle as orc
    ue, obje, ste, polished, slack = orc.qp_solve_exact(oq, x0, goal)
    band = np.abs(slack) <= 1e-6
    status = np.where(res.status == 0, 0, 1)
    assert not (res.status == 2).any(), f"{(res.status == 2).sum()} samples hit max_iter"
    np.testing.assert_array_equal(status[~band], ste[~band])
    ok = (ste == 0) & ~band & polished
    assert ok.sum() >= min_feasible
    du = np.abs(res.u0[ok] - ue[ok, :2]).max()
    assert du <= U_TOL, f"first input differs by {du}"
    rel = np.abs(res.objective[ok] - obje[ok]) / np.maximum(1.0, np.abs(obje[ok]))
    assert rel.max() <= OBJ_RTOL, f"objective differs by {rel.max()} (relative)"
    if res.u_full is not None:
        assert np.abs(res.u_full[ok] - ue[ok]).max() <= U_TOL
    bad = ste == 1
    assert np.all(np.isinf(res.objective[bad & ~band])) and np.all(np.isnan(res.u0[bad & ~band]))
    return int(band.sum()), float(du), float(rel.max())


@pytest.mark.parametrize("env_name,N,n_states", [("RoadOneCarEnv", 20, 400), ("RoadEnv", 20, 200), ("RoadMultipleCarsEnv", 20, 200),
                                                 ("RoadOneCarEnv", 10, 200), ("RoadOneCarEnv", 5, 100), ("RoadOneCarEnv", 1, 100),
                                                 ("RoadOneCarEnv", 40, 120), ("RoadMultipleCarsEnv", 40, 60),
                                                 ("RoadOneCarEnv", 80, 60), ("RoadEnv", 80, 40)])
def test_qp_solutions_match_exact_oracle(torch_cuda, env_name, N, n_states):
    c, bq, oq = _setup(env_name, N)
    x0 = _states(c, n_states, seed=N)
    res = bq.solve_host(x0, want_u_full=True)
    n_band, du, rel = _compare(res, x0, oq, np.array(c.goal, dtype=float))
    assert n_band <= 3
    total_iters, launches = bq.last_stats()
    assert launches >= 2 and total_iters == int(res.iters.sum())


def test_known_answer_single_qp(torch_cuda):
    """a state with saturated acceleration; the solution must satisfy the KKT conditions of the reference's
    QP (lib/mpc.py:321-332) built from the reference-identical matrices."""
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    x0 = np.array([[12.0, 0.4, 0.05, 2.0]])
    res = bq.solve_host(x0, want_u_full=True)
    assert res.status[0] == 0
    u = res.u_full[0]
    assert abs(u[0] - 2.0) < 1e-9                              # a(0) at its bound
    ub = oq.rhs(x0)[0]
    assert (oq.G @ u - ub).max() <= 1e-8
    grad = oq.H @ u + oq.lin(x0, c.goal)[0]
    act = np.flatnonzero(oq.G @ u - ub > -1e-7)
    lam, *_ = np.linalg.lstsq(oq.G[act].T, -grad, rcond=None)
    assert np.abs(oq.G[act].T @ lam + grad).max() < 1e-6 and lam.min() > -1e-7


def test_device_and_host_entry_points_agree_and_are_deterministic(torch_cuda):
    torch = torch_cuda
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    for B in (1, 2, 127, 129, 1000):
        x0 = _states(c, B, seed=B)
        host = bq.solve_host(x0, want_u_full=True)
        dev = bq.solve(torch.from_numpy(np.ascontiguousarray(x0.T)).cuda(), want_u_full=True)
        np.testing.assert_array_equal(dev["status"].cpu().numpy(), host.status)
        np.testing.assert_array_equal(dev["iters"].cpu().numpy(), host.iters)
        np.testing.assert_array_equal(dev["u0"].cpu().numpy().T, host.u0)
        np.testing.assert_array_equal(dev["objective"].cpu().numpy(), host.objective)
        np.testing.assert_array_equal(dev["u_full"].cpu().numpy(), host.u_full)
    # empty batch
    dev = bq.solve(torch.empty((4, 0), dtype=torch.float64, device="cuda"))
    assert dev["status"].numel() == 0


def test_infeasible_and_degenerate_inputs(torch_cuda):
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    g = np.array(c.goal, dtype=float)
    x0 = np.array([
        g,                                   # at the goal: u = 0, nothing active
        [31.0, 1.5, 0.0, 0.0],               # behind the obstacle line x <= 30: infeasible through the u-free rows
        [5.0, 1.5, 0.0, 5.0],                # feasible far away
        [29.0, 1.5, 0.0, 5.0],               # too fast, too close: infeasible
        [np.nan, 0.0, 0.0, 0.0],             # NaN state: infeasible, never a crash
        [10.0, 2.99, 0.39, 3.0],             # heading into the road edge
    ])
    res = bq.solve_host(x0, want_u_full=True)
    assert res.status[0] == 0 and np.abs(res.u_full[0]).max() < 1e-9 and abs(res.objective[0]) < 1e-9
    assert res.status[1] == 1 and res.status[3] == 1 and res.status[4] == 1
    assert res.status[2] == 0
    from oracle import carmpc_oracle as orc
    ue, obje, ste, polished, slack = orc.qp_solve_exact(oq, x0[[0, 1, 2, 3, 5]], g)
    np.testing.assert_array_equal(np.where(res.status[[0, 1, 2, 3, 5]] == 0, 0, 1), ste)


def test_raw_admm_mode_is_close_but_polish_is_what_meets_the_tolerance(torch_cuda):
    """polish=0 returns the float32 ADMM iterate (OSQP-like accuracy); with eps 1e-5 it is within ~1e-3 of the exact
    solution, which is why the default path polishes in float64."""
    from oracle import carmpc_oracle as orc
    c, bq, oq = _setup("RoadOneCarEnv", 20, polish=0, eps_abs=1e-5, eps_rel=1e-5)
    x0 = _states(c, 200, seed=3)
    res = bq.solve_host(x0)
    ue, obje, ste, polished, slack = orc.qp_solve_exact(oq, x0, np.array(c.goal, dtype=float))
    ok = (ste == 0) & (res.status == 0)
    assert ok.sum() > 50
    assert np.abs(res.u0[ok] - ue[ok, :2]).max() < 5e-3


def test_controller_step_is_a_drop_in(torch_cuda):
    """MPCStateFB.step / MPCOutputFB.step keep the reference's contract (lib/mpc.py:312-349, :450-492): return u(0),
    set the side-effect attributes, raise OutsideTheRegionOfAttractionError when infeasible."""
    from carmpc_b200.lib.mpc import OutsideTheRegionOfAttractionError
    from oracle import carmpc_oracle as orc
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    g = np.array(c.goal, dtype=float)
    x0 = np.array([12.0, 0.4, 0.05, 2.0])
    u0 = c.step(x0)
    ue, obje, *_ = orc.qp_solve_exact(oq, x0[None, :], g)
    assert u0.shape == (2,) and np.abs(u0 - ue[0, :2]).max() <= U_TOL
    assert c.x_horizon.shape == (21, 4) and c.u_horizon.shape == (20, 2)
    np.testing.assert_allclose(c.x_horizon[0], x0, atol=1e-12)
    xs = (c.T @ x0 + c.S @ ue[0]).reshape(-1, 4)
    np.testing.assert_allclose(c.x_horizon, xs, atol=1e-3)
    assert abs(c.cost - obje[0]) <= OBJ_RTOL * abs(obje[0])
    assert abs(c.terminal_cost - xs[-1] @ c.P @ xs[-1]) <= 1e-2 * abs(c.terminal_cost)     # absolute x(N): reference quirk
    with pytest.raises(OutsideTheRegionOfAttractionError):
        c.step(np.array([29.0, 1.5, 0.0, 5.0]))
    # output feedback: observer update, then the same QP on the estimate
    ofb = make_controller(make_env("RoadEnv"), 20, cls="MPCOutputFB", init_state=[5, -1.5, 0, 0])
    y = np.array([5.1, -1.45, 0.4])
    xhat = ofb.luenberger_observer(y)
    u = ofb.step(y)
    np.testing.assert_allclose(ofb.x_estimate, xhat, atol=1e-14)
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", FIXTURE["RoadEnv"]))
    oq2 = orc.CondensedQP("RoadEnv", 20, Ab)
    ue2, *_ = orc.qp_solve_exact(oq2, xhat[None, :], np.array([30, 1.5, 0, 0.0]))
    assert np.abs(u - ue2[0, :2]).max() <= U_TOL
    np.testing.assert_array_equal(ofb.previous_u, u)


def test_closed_loop_matches_oracle(torch_cuda):
    """BASELINE config 4 in miniature: output-feedback runs against the nonlinear bicycle, order of
    examples/run_MPCOutputFB.py:29-41.  The oracle steps the same loop with exact QP solutions."""
    torch = torch_cuda
    from oracle import carmpc_oracle as orc
    c, bq, oq = _setup("RoadEnv", 20)
    g = np.array(c.goal, dtype=float)
    rng = np.random.default_rng(0)
    R, steps = 12, 25
    x_init = rng.uniform([0, -2.5, -0.2, 0], [10, 2.5, 0.2, 3], size=(R, 4))

    def exact(x):
        u, obj, st, *_ = orc.qp_solve_exact(oq, x, g)
        return u[:, :2], st

    for output_feedback in (True, False):
        want_x, want_fail, want_traj = orc.closed_loop(oq, x_init, g, steps, output_feedback, exact)
        out = bq.closed_loop(torch.from_numpy(np.ascontiguousarray(x_init.T)).cuda(), steps, c.A, c.B,
                             C=orc.C_OUT if output_feedback else None, L=orc.L_OBS if output_feedback else None,
                             want_traj=True, want_inputs=True)
        fail = out["fail_step"].cpu().numpy()
        np.testing.assert_array_equal(fail, want_fail)
        traj = out["traj"].cpu().numpy().transpose(0, 2, 1)                   # (steps, R, 4)
        alive = want_fail < 0
        assert alive.sum() >= R // 2
        assert np.abs(traj[:, alive] - want_traj[:, alive]).max() <= 1e-6
        np.testing.assert_allclose(out["final"].cpu().numpy().T[alive], want_x[alive], atol=1e-6)
        assert out["total_iters"] > 0


def test_config3_full_grid_properties(torch_cuda):
    """BASELINE config 3: 10^6 initial states on the RoadOneCarEnv grid, one horizon-20 QP each.  The oracle cannot
    solve 10^6 QPs in seconds, so: every solved sample is polished (KKT-certified in float64 on the device),
    a strided sample of 600 is compared with the exact oracle, feasibility is monotone along rays towards the goal
    (convexity of the feasible set), and a second solve is bit-identical (idempotence)."""
    torch = torch_cuda
    from carmpc_b200.grids import config3_axes, materialise_grid
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    axes = config3_axes()
    x0 = torch.stack(materialise_grid(axes, device="cuda")).contiguous()          # (4, 10^6)
    n = x0.shape[1]
    assert n == 10 ** 6
    out = bq.solve(x0)
    status = out["status"].cpu().numpy()
    assert (status == 2).sum() == 0
    frac = (status == 0).mean()
    assert 0.3 < frac < 0.95
    out2 = bq.solve(x0)
    assert torch.equal(out["status"], out2["status"]) and torch.equal(out["iters"], out2["iters"])
    assert torch.equal(out["u0"].nan_to_num(7.0), out2["u0"].nan_to_num(7.0))
    idx = np.arange(0, n, n // 600)[:600]
    from oracle import carmpc_oracle as orc
    xs = x0[:, torch.from_numpy(idx).cuda()].cpu().numpy().T
    ue, obje, ste, polished, slack = orc.qp_solve_exact(oq, xs, np.array(c.goal, dtype=float))
    band = np.abs(slack) <= 1e-6
    np.testing.assert_array_equal(np.where(status[idx] == 0, 0, 1)[~band], ste[~band])
    ok = (ste == 0) & ~band & polished
    u0 = out["u0"].cpu().numpy().T[idx]
    assert np.abs(u0[ok] - ue[ok, :2]).max() <= U_TOL
    obj = out["objective"].cpu().numpy()[idx]
    assert (np.abs(obj[ok] - obje[ok]) / np.maximum(1, np.abs(obje[ok]))).max() <= OBJ_RTOL
    # convexity of the feasible set: the midpoint of two feasible states is feasible
    feas = np.flatnonzero(status == 0)
    rng = np.random.default_rng(0)
    a, b = rng.choice(feas, 4000), rng.choice(feas, 4000)
    xa = x0[:, torch.from_numpy(a).cuda()]
    xb = x0[:, torch.from_numpy(b).cuda()]
    mid = bq.solve((0.5 * (xa + xb)).contiguous())
    assert int((mid["status"] != 0).sum()) == 0


def test_disturbance_controller_shifts_the_constraints(torch_cuda):
    """MPCOutputFBWithDisturbance (lib/mpc.py:495-667, experimental in the reference): the scalar disturbance estimate
    enters the prediction as x_ = T x0 + S u + ABd d, i.e. a per-sample shift of the constraint right-hand sides.
    Checked against the exact oracle on the shifted problem."""
    from oracle import carmpc_oracle as orc
    from carmpc_b200.batch import BatchQP
    env = make_env("RoadEnv")
    ctl = make_controller(env, 20, cls="MPCOutputFBWithDisturbance", init_state=[20, 0.5, 0, 2])
    bq = BatchQP.from_controller(ctl)
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", FIXTURE["RoadEnv"]))
    oq = orc.CondensedQP("RoadEnv", 20, Ab)
    ABd = ctl.disturbance_response()
    x_ref = np.array([30, 1.5, 0, 0.0])
    rng = np.random.default_rng(8)
    x0 = x_ref + rng.uniform(-1, 1, size=(60, 4)) * np.array([10.0, 1.0, 0.2, 2.0])
    d = rng.uniform(-0.02, 0.02, size=60)
    res = bq.solve_host(x0, x_ref=x_ref, want_u_full=True, c=d)
    # oracle: same QP with the bounds shifted by rows_x @ ABd * d  (one-sided form: G u <= w - Gx x0 - Gd d)
    rows = np.vstack((orc.terminal_constraint(Ab, 20)[0], np.zeros((80, 84)), orc.state_constraint("RoadEnv", 20)[0]))
    Gd = rows @ ABd
    n_ok = 0
    for i in range(len(x0)):
        class Shifted(orc.CondensedQP):
            pass
        q = Shifted("RoadEnv", 20, Ab)
        q.w = q.w - Gd * d[i]
        ue, obje, ste, pol, slack = orc.qp_solve_exact(q, x0[i:i + 1], x_ref)
        if abs(slack[0]) <= 1e-6:
            continue
        assert (res.status[i] == 0) == (ste[0] == 0)
        if ste[0] == 0 and pol[0]:
            n_ok += 1
            assert np.abs(res.u_full[i] - ue[0]).max() <= U_TOL
            assert abs(res.objective[i] - obje[0]) <= OBJ_RTOL * max(1.0, abs(obje[0]))
    assert n_ok >= 10
    # and the drop-in step() runs end to end (observer with disturbance state, then the QP)
    u = ctl.step(np.array([20.0, 0.5, 2.0]))     # innovation 0: the (experimental) disturbance estimate stays 0
    assert u.shape == (2,) and np.all(np.isfinite(u))


def test_config4_monte_carlo_properties(torch_cuda):
    """BASELINE config 4 at full size: 10^5 output-feedback closed loops x 200 steps against the nonlinear bicycle.
    Too large for the oracle; properties instead: bit-identical when repeated, a run that never failed reaches the
    goal position and speed, a failed run keeps the state it had when its QP became infeasible, inputs within the
    actuator box."""
    torch = torch_cuda
    from carmpc_b200.batch import BatchQP
    from carmpc_b200.lib.mpc import _C_XYV, _L_OBSERVER
    ctl = make_controller(make_env("RoadEnv"), 20)
    bq = BatchQP.from_controller(ctl)
    R, T = 100_000, 200
    g = torch.Generator(device="cpu").manual_seed(0)
    lo = torch.tensor([0.0, -2.5, -0.2, 0.0], dtype=torch.float64)
    hi = torch.tensor([10.0, 2.5, 0.2, 3.0], dtype=torch.float64)
    x_init = (lo[:, None] + (hi - lo)[:, None] * torch.rand((4, R), generator=g, dtype=torch.float64)).cuda().contiguous()
    a = bq.closed_loop(x_init, T, ctl.A, ctl.B, C=_C_XYV, L=_L_OBSERVER)
    b = bq.closed_loop(x_init, T, ctl.A, ctl.B, C=_C_XYV, L=_L_OBSERVER)
    assert torch.equal(a["fail_step"], b["fail_step"]) and torch.equal(a["final"], b["final"])
    fail = a["fail_step"]
    ok = fail < 0
    assert 0.05 < ok.float().mean().item() < 0.95
    goal = torch.tensor([30.0, 1.5, 0.0, 0.0], dtype=torch.float64, device="cuda")
    err = (a["final"] - goal[:, None]).abs()
    # x and v converge to the goal; y and psi need not (at v -> 0 the bicycle has no steering authority), but every
    # surviving run stays on the road and inside the heading bound
    worst = err[:, ok].max(dim=1).values
    assert worst[0] <= 1e-3 and worst[3] <= 1e-3
    assert (a["final"][1, ok].abs() <= 3.0 + 1e-6).all() and (a["final"][2, ok].abs() <= np.pi / 8 + 1e-3).all()
    # a short rerun with logs: inputs respect the input box, failed runs freeze
    small = bq.closed_loop(x_init[:, :2000].contiguous(), 60, ctl.A, ctl.B, C=_C_XYV, L=_L_OBSERVER, want_traj=True,
                           want_inputs=True)
    u = small["inputs"]
    assert (u[:, 0].abs() <= 2.0 + 1e-9).all() and (u[:, 1].abs() <= np.pi / 8 + 1e-9).all()
    f = small["fail_step"].cpu().numpy()
    traj = small["traj"].cpu().numpy()
    for r in np.flatnonzero(f >= 0)[:50]:
        assert np.array_equal(traj[f[r]:, :, r], np.repeat(traj[f[r]:f[r] + 1, :, r], 60 - f[r], axis=0))


@pytest.mark.parametrize("terminal,inputs,state", [(False, True, False), (False, False, False), (True, False, True),
                                                   (False, True, True), (True, True, False)])
def test_constraint_block_switches(torch_cuda, terminal, inputs, state):
    """The reference's constructor flags (terminal_constraint / input_constraint / state_constraint, lib/mpc.py:22-31)
    switch whole constraint blocks off; the condensed problem then has no general rows, no box, or neither."""
    from oracle import carmpc_oracle as orc
    from carmpc_b200.batch import BatchQP
    env = make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0])
    c = make_controller(env, 10, terminal_constraint=terminal, input_constraint=inputs, state_constraint=state)
    bq = BatchQP.from_controller(c)
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", FIXTURE["RoadOneCarEnv"]))
    oq = orc.CondensedQP("RoadOneCarEnv", 10, Ab, use_terminal=terminal, use_input=inputs, use_state=state)
    g = np.array(c.goal, dtype=float)
    x0 = _states(c, 120, seed=5, spread=(6.0, 1.0, 0.2, 1.5))
    res = bq.solve_host(x0, want_u_full=True)
    if not (terminal or state):
        assert (res.status == 0).all()                     # nothing but (at most) the input box: always feasible
    if not (terminal or inputs or state):
        want = -np.linalg.solve(oq.H, oq.lin(x0, g).T).T    # unconstrained optimum
        np.testing.assert_allclose(res.u_full, want, atol=1e-8)
        return
    ue, obje, ste, polished, slack = orc.qp_solve_exact(oq, x0, g)
    ok = (ste == 0) & polished & (np.abs(slack) > 1e-6) & (res.status == 0)
    assert ok.sum() >= 20
    assert np.abs(res.u_full[ok] - ue[ok]).max() <= U_TOL
    assert (np.abs(res.objective[ok] - obje[ok]) / np.maximum(1, np.abs(obje[ok]))).max() <= OBJ_RTOL
    if inputs:                                             # with a finite input box the flags are exact
        band = np.abs(slack) <= 1e-6
        np.testing.assert_array_equal(np.where(res.status == 0, 0, 1)[~band], ste[~band])
    else:                                                  # without it a feasible state is never called infeasible
        assert not ((res.status == 1) & (ste == 0)).any()


# ----------------------------------------------------------------------------------------------------------------
# seeded solve (region-of-attraction maps): same certified optima and flags as the cold solve, whatever the seeds
# ----------------------------------------------------------------------------------------------------------------
def _grid_states(torch, axes):
    from carmpc_b200.grids import materialise_grid
    return torch.stack(materialise_grid(axes, device="cuda")).contiguous()


def _assert_same_solution(a, b, oq=None, x0=None):
    """Flags must be identical; the only exemption is BASELINE's: states whose exact feasibility slack (oracle LP) is
    within 1e-6 of zero, which are enumerated here (the two paths may stop on different sides of such a boundary)."""
    sa, sb = a["status"].cpu().numpy(), b["status"].cpu().numpy()
    diff = np.flatnonzero(sa != sb)
    if len(diff):
        assert oq is not None and len(diff) <= 3, f"{len(diff)} feasibility flags differ"
        from oracle import carmpc_oracle as orc
        _, slack = orc.qp_feasible_lp(oq, x0.cpu().numpy().T[diff])
        assert np.all(np.abs(slack) <= 1e-6), f"flags differ outside the boundary band: slack {slack}"
    ok = (sa == 0) & (sb == 0)
    sa = np.where(sa == sb, sa, -1)
    ua, ub = a["u0"].cpu().numpy()[:, ok], b["u0"].cpu().numpy()[:, ok]
    assert np.abs(ua - ub).max() <= 1e-7                      # two certified KKT points of a strictly convex QP
    oa, ob = a["objective"].cpu().numpy(), b["objective"].cpu().numpy()
    assert np.abs(oa[ok] - ob[ok]).max() <= 1e-8 * max(1.0, np.abs(oa[ok]).max())
    assert np.all(np.isinf(ob[sa == 1])) and np.all(np.isnan(b["u0"].cpu().numpy()[:, sa == 1]))
    if "u_full" in a:
        assert np.abs(a["u_full"].cpu().numpy()[ok] - b["u_full"].cpu().numpy()[ok]).max() <= 1e-7


@pytest.mark.parametrize("env_name,N", [("RoadOneCarEnv", 20), ("RoadMultipleCarsEnv", 10), ("RoadEnv", 40)])
def test_seeded_grid_solve_matches_cold_solve_and_oracle(torch_cuda, env_name, N):
    torch = torch_cuda
    from carmpc_b200.grids import lattice_seeds
    c, bq, oq = _setup(env_name, N)
    g = np.array(c.goal, dtype=float)
    axes = [np.linspace(g[0] - 20.0, g[0] + 0.5, 24), np.linspace(-3.0, 3.0, 40), np.linspace(-0.3, 0.3, 3),
            np.linspace(-1.0, 4.0, 4)]
    x0 = _grid_states(torch, axes)
    B = x0.shape[1]
    cold = bq.solve(x0, want_u_full=True)
    seed = torch.from_numpy(lattice_seeds([len(a) for a in axes], block=(2, 8, 1, 1))).cuda()
    warm = bq.solve(x0, want_u_full=True, seed=seed)
    _assert_same_solution(cold, warm, oq, x0)
    st = cold["status"].cpu().numpy()
    assert (st == 0).sum() > B // 10 and (st == 1).sum() > 0, "the grid should cross the region-of-attraction boundary"
    n_anchor = int((seed.cpu().numpy() == np.arange(B)).sum())
    assert warm["seeded"] >= ((st == 0).sum() - n_anchor) // 2, "most feasible followers should certify from their seed"
    assert int(warm["iters"].sum().item()) < int(cold["iters"].sum().item())
    # oracle spot check of the seeded result (BASELINE tolerances)
    from carmpc_b200.batch import QPResult
    pick = np.random.default_rng(1).choice(B, size=60, replace=False)
    res = QPResult(u0=warm["u0"].cpu().numpy().T[pick], objective=warm["objective"].cpu().numpy()[pick],
                   status=st[pick], iters=warm["iters"].cpu().numpy()[pick], u_full=warm["u_full"].cpu().numpy()[pick])
    _compare(res, x0.cpu().numpy().T[pick], oq, g, min_feasible=3)


def test_seeded_solve_is_independent_of_the_seeds(torch_cuda):
    """Arbitrary seed maps (random anchors far away, chains, out-of-range entries, all-anchor, single anchor)."""
    torch = torch_cuda
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    x0 = torch.from_numpy(np.ascontiguousarray(_states(c, 3000, seed=7).T)).cuda()
    B = x0.shape[1]
    cold = bq.solve(x0)
    rng = np.random.default_rng(3)
    ident = np.arange(B, dtype=np.int32)
    maps = {"identity": ident, "single anchor": np.zeros(B, dtype=np.int32),
            "random": rng.integers(0, B, size=B).astype(np.int32),
            "out of range": np.where(rng.random(B) < 0.5, -1, B + 5).astype(np.int32),
            "chain": np.maximum(ident - 1, 0).astype(np.int32)}
    for name, m in maps.items():
        out = bq.solve(x0, seed=torch.from_numpy(m).cuda())
        _assert_same_solution(cold, out, oq, x0)
    # the empty batch
    out = bq.solve(torch.empty((4, 0), dtype=torch.float64, device="cuda"), seed=torch.empty(0, dtype=torch.int32, device="cuda"))
    assert out["status"].numel() == 0 and out["seeded"] == 0


def test_map_host_entry_point_matches_the_batched_solve(torch_cuda):
    """carmpc_qp_map_host (grid axes in, numpy out): cold and seeded maps against the batched solve of the materialised grid,
    including a permuted axis order (the state of axis k is axis_to_state[k])."""
    torch = torch_cuda
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    axes = [np.linspace(8.0, 30.4, 18), np.linspace(-2.8, 2.9, 33), np.linspace(-0.3, 0.3, 3), np.linspace(-1.0, 4.0, 4)]
    x0 = _grid_states(torch, axes)
    ref = bq.solve(x0)
    st, u0, obj = ref["status"].cpu().numpy(), ref["u0"].cpu().numpy().T, ref["objective"].cpu().numpy()
    ok = st == 0
    for block in (None, (3, 8, 1, 1), (2, 5, 3, 2)):
        res = bq.solve_map_host(axes, block=block)
        np.testing.assert_array_equal(res.status, st)
        assert np.abs(res.u0[ok] - u0[ok]).max() <= 1e-7 and np.abs(res.objective[ok] - obj[ok]).max() <= 1e-7
        assert np.all(np.isnan(res.u0[~ok])) and np.all(np.isinf(res.objective[~ok]))
        assert (res.seeded > 0) == (block is not None)
    # axes given as (v, y, x, psi): same states in another order
    perm = (3, 1, 0, 2)
    res = bq.solve_map_host([axes[k] for k in perm], block=(1, 8, 3, 1), axis_to_state=perm, want_objective=False)
    dims = [len(axes[k]) for k in perm]
    idx = np.stack(np.unravel_index(np.arange(int(np.prod(dims))), dims), axis=0)      # index along each given axis
    flat = np.ravel_multi_index([idx[perm.index(s)] for s in range(4)], [len(a) for a in axes])
    np.testing.assert_array_equal(res.status, st[flat])
    assert res.objective is None and np.abs(res.u0[ok[flat]] - u0[flat][ok[flat]]).max() <= 1e-7
    with pytest.raises(Exception):
        bq.solve_map_host(axes, axis_to_state=(0, 1, 1, 3))


def test_pinned_host_results_equal_pageable_ones(torch_cuda):
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    x0 = _states(c, 500, seed=11)
    a = bq.solve_host(x0, want_u_full=True)
    b = bq.solve_host(x0, want_u_full=True, pinned=True)
    for name in ("u0", "objective", "status", "iters", "u_full"):
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name))
    keep = b.u0.copy()
    b2 = bq.solve_host(x0[::-1].copy(), pinned=True)                  # same size: the pinned buffers are reused
    assert b2.u0.ctypes.data == b.u0.ctypes.data and b2.u_full is None
    np.testing.assert_array_equal(b2.u0[::-1], keep)


def test_seeded_solve_with_per_sample_disturbance(torch_cuda):
    """With a per-sample disturbance the anchors export no multiplier maps / certificates (they assume a common shift),
    but the seeded solve must still reproduce the cold one: anchor's active set -> polish -> ADMM for the rest."""
    torch = torch_cuda
    from carmpc_b200.batch import BatchQP
    from carmpc_b200.grids import lattice_seeds
    ctl = make_controller(make_env("RoadEnv"), 20, cls="MPCOutputFBWithDisturbance", init_state=[20, 0.5, 0, 2])
    bq = BatchQP.from_controller(ctl)
    x_ref = np.array([30, 1.5, 0, 0.0])
    axes = [np.linspace(12.0, 30.5, 20), np.linspace(-2.6, 2.6, 24), np.linspace(-0.2, 0.2, 3), np.linspace(0.0, 4.0, 3)]
    x0 = _grid_states(torch, axes)
    B = x0.shape[1]
    d = torch.from_numpy(np.random.default_rng(5).uniform(-0.02, 0.02, size=B)).cuda()
    cold = bq.solve(x0, x_ref=x_ref, c=d)
    seed = torch.from_numpy(lattice_seeds([len(a) for a in axes], block=(2, 6, 1, 1))).cuda()
    warm = bq.solve(x0, x_ref=x_ref, c=d, seed=seed)
    _assert_same_solution(cold, warm)
    st = cold["status"].cpu().numpy()
    assert (st == 0).sum() > B // 10 and warm["seeded"] > 0
    assert bq.polish_stats()["used_multiplier_map"] == 0


def test_seeded_entry_points_reject_what_they_cannot_do(torch_cuda):
    """Error behaviour of the seeded / map entry points: no float64 polish -> no seeded solve; malformed arguments."""
    torch = torch_cuda
    from carmpc_b200._capi import CarmpcError
    c, bq, oq = _setup("RoadOneCarEnv", 20, polish=0)
    x0 = torch.from_numpy(np.ascontiguousarray(_states(c, 64, seed=1).T)).cuda()
    seed = torch.arange(64, dtype=torch.int32, device="cuda")
    with pytest.raises(CarmpcError, match="polish"):
        bq.solve(x0, seed=seed)
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    with pytest.raises(ValueError):
        bq.solve(x0, seed=seed[:10])
    with pytest.raises(ValueError):
        bq.solve(x0, seed=seed.to(torch.int64))
    with pytest.raises(CarmpcError, match="axis"):
        bq.solve_map_host([np.zeros(5000), [0.0], [0.0], [1.0]])
    with pytest.raises(CarmpcError, match="block"):
        bq.solve_map_host([np.linspace(20, 30, 4), [1.0], [0.0], [1.0]], block=(0, 1, 1, 1))
    res = bq.solve_map_host([np.linspace(20, 30, 4), [1.0], [0.0], [1.0]], block=(2, 1, 1, 1))
    assert res.status.shape == (4,) and res.u0.shape == (4, 2)


# ----------------------------------------------------------------------------------------------------------------
# parity on the BASELINE configurations themselves ("close the parity holes")
# ----------------------------------------------------------------------------------------------------------------
def _config_states(ref, n_states, seed):
    """Half of the states from the config-3/5 region-of-attraction grid (strided), half uniform around the goal."""
    from carmpc_b200.grids import config3_axes, materialise_grid_host, grid_size
    axes = config3_axes()
    n = grid_size(axes)
    k = n_states // 2
    idx = (np.arange(k, dtype=np.int64) * (n // k) + seed * 7919) % n
    dims = [len(a) for a in axes]
    sub = np.stack([a[i] for a, i in zip(axes, np.unravel_index(idx, dims))], axis=1)
    sub[:, 0] += ref.goal[0] - 29.9                       # the grid is laid out for the RoadOneCarEnv goal
    rng = np.random.default_rng(seed)
    rnd = ref.goal + rng.uniform(-1, 1, size=(n_states - k, 4)) * np.array([12.0, 1.4, 0.25, 2.5])
    return np.vstack((sub, rnd))


@pytest.mark.parametrize("env_name,N,n_states,n_lp", [("RoadOneCarEnv", 20, 10000, 300), ("RoadEnv", 20, 10000, 300),
                                                      ("RoadMultipleCarsEnv", 20, 10000, 300), ("RoadOneCarEnv", 10, 10000, 300),
                                                      ("RoadOneCarEnv", 40, 4000, 100), ("RoadOneCarEnv", 80, 2000, 40)])
def test_gpu_results_are_kkt_points_of_the_reference_qp(torch_cuda, env_name, N, n_states, n_lp):
    """No oracle solver in the loop: the QP is assembled from the matrices the unmodified reference controller produced
    (tests/golden/qp_*.npz, lib/mpc.py:318-332) and every GPU result must satisfy its optimality conditions in float64
    (primal residual <= 1e-8, stationarity <= 1e-6 with non-negative multipliers on the active rows); a sample of the
    'infeasible' flags is confirmed by a phase-1 LP on the same rows."""
    from kkt_check import ReferenceQP, assert_results_satisfy_reference_qp
    ref = ReferenceQP(env_name, N)
    c, bq, _ = _setup(env_name, N)
    np.testing.assert_array_equal(np.array(c.goal, dtype=float), ref.goal)
    x0 = _config_states(ref, n_states, seed=N)
    res = bq.solve_host(x0, want_u_full=True)
    info = assert_results_satisfy_reference_qp(ref, x0, res.u_full, res.status, objective=res.objective, n_lp=n_lp, seed=N)
    assert info["solved"] >= n_states // 10 and info["infeasible"] >= n_states // 50, info
    np.testing.assert_array_equal(res.u0[res.status == 0], res.u_full[res.status == 0][:, :2])
    print(f"{env_name} N={N}: {info}")


@pytest.mark.parametrize("N,n_states", [(10, 400), (40, 320), (80, 300)])
def test_config5_grid_states_match_exact_oracle(torch_cuda, N, n_states):
    """BASELINE config 5: strided states OF THE CONFIG GRID at the other horizons of the sweep, against the exact
    oracle at BASELINE's tolerances (N = 20 is covered by test_config3_full_grid_properties)."""
    torch = torch_cuda
    from carmpc_b200.grids import config3_axes, materialise_grid, grid_size
    c, bq, oq = _setup("RoadOneCarEnv", N)
    axes = config3_axes()
    n = grid_size(axes)
    x0 = torch.stack(materialise_grid(axes, device="cuda")).contiguous()
    out = bq.solve(x0)
    status = out["status"].cpu().numpy()
    assert (status == 2).sum() <= 2, "undecided (max_iter) states on the config grid"
    idx = np.arange(0, n, n // n_states)[:n_states] + (N % 7) * 13
    xs = x0[:, torch.from_numpy(idx).cuda()].cpu().numpy().T
    from carmpc_b200.batch import QPResult
    res = QPResult(u0=out["u0"].cpu().numpy().T[idx], objective=out["objective"].cpu().numpy()[idx], status=status[idx],
                   iters=out["iters"].cpu().numpy()[idx])
    n_band, du, rel = _compare(res, xs, oq, np.array(c.goal, dtype=float), min_feasible=n_states // 10)
    assert n_band <= 3


def test_config4_closed_loops_match_the_oracle_fixture(torch_cuda):
    """BASELINE config 4 in its own shape: 200 runs x 200 steps, output feedback and state feedback, against the
    oracle's closed loop with exact QP solutions (tests/golden/closed_loop_config4.npz, written by
    gen_oracle_fixtures.py; tests/test_oracle_golden.py re-derives part of it live): fail_step equal, trajectories
    and final states within 1e-6."""
    torch = torch_cuda
    from carmpc_b200.lib.mpc import _C_XYV, _L_OBSERVER
    fx = np.load(os.path.join(GOLDEN, "closed_loop_config4.npz"))
    c, bq, _ = _setup("RoadEnv", 20)
    x_init = fx["x_init"]
    steps, stride = int(fx["steps"]), int(fx["stride"])
    for tag, fb in (("ofb", True), ("sfb", False)):
        out = bq.closed_loop(torch.from_numpy(np.ascontiguousarray(x_init.T)).cuda(), steps, c.A, c.B,
                             C=_C_XYV if fb else None, L=_L_OBSERVER if fb else None, want_traj=True)
        fail, want_fail = out["fail_step"].cpu().numpy(), fx[f"fail_{tag}"]
        np.testing.assert_array_equal(fail, want_fail)
        alive = want_fail < 0
        assert alive.sum() >= len(alive) // 4
        traj = out["traj"].cpu().numpy().transpose(0, 2, 1)[stride - 1::stride]
        assert np.abs(traj[:, alive] - fx[f"traj_{tag}"][:, alive]).max() <= 1e-6
        np.testing.assert_allclose(out["final"].cpu().numpy().T, fx[f"final_{tag}"], atol=1e-6)
