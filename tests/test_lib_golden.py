"""The host-side mirror of the reference interface (carmpc_b200.lib) against golden vectors from the reference."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, TERMINAL_SETS, FIXTURES, golden, make_env, make_controller


def test_model_and_gains():
    g = golden("model.npz")
    c = make_controller(make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0]))
    np.testing.assert_array_equal(c.A, g["A"])
    np.testing.assert_array_equal(c.B, g["B"])
    np.testing.assert_allclose(c.P, g["P"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(c.K, g["K"], rtol=0, atol=1e-11)
    np.testing.assert_array_equal(c.input_upper, g["input_upper"])
    np.testing.assert_array_equal(c.input_lower, g["input_lower"])
    ofb = make_controller(make_env("RoadEnv"), cls="MPCOutputFB", init_state=[5, -1.5, 0, 0])
    np.testing.assert_array_equal(ofb.L, g["L"])
    np.testing.assert_array_equal(ofb.C, g["C"])
    np.testing.assert_array_equal(ofb.y_goal, g["y_goal"])


@pytest.mark.parametrize("N", [1, 5, 10, 20])
def test_predmod_costgen(N):
    from carmpc_b200.lib.matrix_gen import predmod, costgen, stack_matrix_along_diag
    g = golden(f"predmod_N{N}.npz")
    m = golden("model.npz")
    T, S = predmod(m["A"], m["B"], N)
    H, h, const = costgen(m["Q"], m["R"], m["P"], T, S, 4)
    np.testing.assert_allclose(T, g["T"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(S, g["S"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(H, g["H"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(h, g["h"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(const, g["const"], rtol=1e-12, atol=1e-9)
    D = stack_matrix_along_diag(m["R"], 3)
    assert D.shape == (6, 6) and np.array_equal(D[2:4, 2:4], m["R"]) and D[0, 2] == 0


@pytest.mark.parametrize("tag,env,goal", [("RoadEnv", "RoadEnv", None), ("RoadOneCarEnv", "RoadOneCarEnv", [29.9, 1.5, 0, 0]),
                                          ("RoadOneCarEnvDefault", "RoadOneCarEnv", None),
                                          ("RoadMultipleCarsEnv", "RoadMultipleCarsEnv", None)])
@pytest.mark.parametrize("N", [1, 3, 20])
def test_constraint_stacks(tag, env, goal, N):
    g = golden(f"constraints_{tag}_N{N}.npz")
    e = make_env(env, goal)
    c = make_controller(e, N)
    for (A, b), (ka, kb) in zip((c.terminal_constraint(), c.input_constraint(), c.state_constraint()),
                                (("At", "bt"), ("Ai", "bi"), ("As", "bs"))):
        np.testing.assert_array_equal(A, g[ka])
        np.testing.assert_array_equal(b, g[kb])
    np.testing.assert_array_equal(np.array(e.goal, dtype=float), g["goal"])
    np.testing.assert_array_equal(np.array(e.constraints_A, dtype=float), g["env_A"])
    np.testing.assert_array_equal(np.array(e.constraints_b, dtype=float), g["env_b"])


def test_missing_terminal_set_falls_back_to_goal_box():
    g = golden("constraints_fallback_N4.npz")
    c = make_controller(make_env("RoadEnv", [10, 0, 0, 0]), 4)
    At, bt = c.terminal_constraint()
    np.testing.assert_array_equal(At, g["At"])
    np.testing.assert_array_equal(bt, g["bt"])


def test_simulator():
    from carmpc_b200.lib.simulator import CarSimulator
    g = golden("simulator.npz")
    m = golden("model.npz")
    for k, clip in ((0, False), (1, True)):
        sim = CarSimulator(dt=float(g[f"dt_{k}"]), clip=clip, C=m["C"])
        sim.reset(g[f"x0_{k}"].copy())
        for u, xs, ys in zip(g[f"u_{k}"], g[f"states_{k}"], g[f"outputs_{k}"]):
            log = sim.step(u)
            np.testing.assert_allclose(sim.state, xs, rtol=0, atol=1e-13)
            np.testing.assert_allclose(sim.output, ys, rtol=0, atol=1e-13)
            assert set(log) == {"car", "inputs"}
        assert abs(sim.time - float(g[f"time_{k}"])) < 1e-12
    sim = CarSimulator(dt=0.2)
    with pytest.raises(AssertionError):
        sim.step([2.5, 0.0])


def test_observer_and_lqr():
    g = golden("observer.npz")
    ofb = make_controller(make_env("RoadEnv"), cls="MPCOutputFB", init_state=[5, -1.5, 0, 0])
    np.testing.assert_array_equal(np.array(ofb.x_estimate, dtype=float), g["xhat"][0])
    for y, u, want in zip(g["y"], g["u"], g["xhat"][1:]):
        ofb.previous_u = u
        ofb.x_estimate = ofb.luenberger_observer(y)
        np.testing.assert_allclose(ofb.x_estimate, want, rtol=0, atol=1e-12)
    x_ref, u_ref = ofb.optimal_target_selection()
    np.testing.assert_allclose(x_ref, [30, 1.5, 0, 0], atol=1e-12)       # unique solution
    np.testing.assert_allclose(u_ref, [0, 0], atol=1e-12)
    gl = golden("lqr.npz")
    lq = make_controller(make_env("RoadMultipleCarsEnv"), cls="MPC", use_LQR=True)
    for x, u, cost in zip(gl["x"], gl["u"], gl["stage_cost"]):
        np.testing.assert_allclose(lq.step(x), u, rtol=0, atol=1e-10)
        np.testing.assert_allclose(lq.stage_cost, cost, rtol=1e-10)


def test_set_goal_requires_equilibrium():
    c = make_controller(make_env("RoadEnv"))
    c.set_goal([10, 1, 0, 0])
    with pytest.raises(AssertionError):
        c.set_goal([10, 1, 0.1, 0])


def test_fourier_motzkin():
    from carmpc_b200.lib.in_adm_set import algorithm_1, algorithm_2
    g = golden("in_adm_set.npz")
    P1, g1 = algorithm_1(g["G"], g["H"][:, 0].copy(), g["phi"])
    P2, g2 = algorithm_2(g["G"], g["H"], g["phi"])
    np.testing.assert_allclose(P1, g["P1"], atol=1e-12)
    np.testing.assert_allclose(g1, g["g1"], atol=1e-12)
    np.testing.assert_allclose(P2, g["P2"], atol=1e-12)
    np.testing.assert_allclose(g2, g["g2"], atol=1e-12)


@pytest.mark.parametrize("file", sorted(FIXTURES))
def test_calc_terminal_set_regenerates_shipped_fixture(file, tmp_path, monkeypatch):
    """calc_terminal_set (Gilbert-Tan with scipy LPs + own polytope algebra) reproduces the reference's shipped
    H-rep row for row (<= 6.3e-13), and writes the same .npy layout."""
    from carmpc_b200.lib import terminal_set as ts
    env_name, goal = FIXTURES[file]
    env = make_env(env_name, goal)
    monkeypatch.setattr(ts, "TERMINAL_SET_DIR", str(tmp_path))
    A, b = ts.calc_terminal_set(env, save=True)
    want = np.load(os.path.join(GOLDEN, "terminal_sets", file))
    assert A.shape == want[:, :4].shape
    np.testing.assert_allclose(np.hstack((A, b[:, None])), want, rtol=0, atol=1e-9)
    written = np.load(os.path.join(str(tmp_path), file))
    assert written.dtype == np.float64 and written.shape == want.shape and written.flags.c_contiguous
    np.testing.assert_allclose(written, want, rtol=0, atol=1e-9)
    # the copy the package ships is the reference's file, bit for bit
    np.testing.assert_array_equal(np.load(os.path.join(TERMINAL_SETS, file)), want)


def test_condensed_qp_is_equivalent_to_reference_rows():
    """carmpc_b200.condensed folds opposite rows and moves u-independent rows to a pre-check; the feasible set in u
    must be unchanged."""
    from carmpc_b200.condensed import build_parametric_qp
    c = make_controller(make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0]))
    pq = build_parametric_qp(c)
    raw = build_parametric_qp(c, merge_rows=False)
    assert pq.m < raw.m and pq.n == 40
    rng = np.random.default_rng(3)
    x0 = rng.uniform([5, -3, -0.39, -1], [30, 3, 0.39, 5], size=(50, 4))
    u = rng.uniform(-1, 1, size=(50, 40)) * np.tile([2.2, 0.42], 20)
    for xi, ui in zip(x0, u):
        xs = c.T @ xi + c.S @ ui
        ok_ref = np.all(pq.rows_x @ xs <= pq.rows_b)
        gu = pq.G @ ui + pq.Gx @ xi
        pre = pq.Px @ xi
        ok_new = np.all(gu <= pq.hi) and np.all(gu >= pq.lo) and np.all(pre <= pq.pre_hi) and np.all(pre >= pq.pre_lo)
        assert ok_ref == ok_new


@pytest.mark.parametrize("N", [1, 5, 20])
def test_disturbance_response_matches_reference_expression(N):
    """``MPCOutputFBWithDisturbance.disturbance_response()`` against ABd as the reference's own expression computes it
    (lib/mpc.py:631-635, recorded by gen_golden.py), and the disturbance model constants (:536-549)."""
    g = golden("disturbance.npz")
    ctl = make_controller(make_env("RoadEnv"), N, cls="MPCOutputFBWithDisturbance", init_state=[20, 0.5, 0, 2])
    np.testing.assert_allclose(ctl.disturbance_response(), g[f"ABd_N{N}"], rtol=0, atol=1e-12 * max(1.0, np.abs(g[f"ABd_N{N}"]).max()))
    np.testing.assert_array_equal(ctl.Bd, g["Bd"])
    np.testing.assert_array_equal(ctl.Cd, g["Cd"])
    np.testing.assert_array_equal(ctl.L1, g["L1"])
    np.testing.assert_array_equal(ctl.L2, g["L2"])
    np.testing.assert_array_equal(ctl.C, g["C"])
