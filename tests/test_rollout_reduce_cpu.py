"""Row reduction of the rollout-form screen (carmpc_b200.batch.RolloutEvaluator.irredundant_rows): the kept rows describe the
same set as the 140 expanded rows a_r A_k^t (what lib/terminal_set.py:203 does with polytope.reduce), and every dropped row
comes with a dual certificate the library can check without an LP solver."""
import numpy as np
import pytest

from conftest import make_env, K_STAR


def _expanded_rows(env, k_steps):
    from carmpc_b200.lib.terminal_set import lqr_closed_loop
    from carmpc_b200.lib import polytope_ops as pc
    _, A_k, A_con, b_con, A_in, b_in = lqr_closed_loop(env)
    goal = np.array(env.goal, dtype=float)
    ps, pi = pc.Polytope(A_con, b_con).translation(-goal), pc.Polytope(A_in, b_in)
    rows, M = [], np.eye(4)
    for t in range(k_steps + 1):
        for a, b in list(zip(ps.A, ps.b)) + (list(zip(pi.A, pi.b)) if t == 0 else []):
            g = a @ M
            rows.append(list(g) + [b + g @ goal])
        M = A_k @ M
    return np.array(rows)


@pytest.mark.parametrize("env_name,goal", [("RoadMultipleCarsEnv", [30, 1.5, 0, 0]), ("RoadOneCarEnv", [29.9, 1.5, 0, 0])])
def test_kept_rows_and_certificates(env_name, goal):
    from carmpc_b200.batch import RolloutEvaluator
    env = make_env(env_name, goal)
    rows = _expanded_rows(env, K_STAR[env_name])
    kept, idx, w = RolloutEvaluator.irredundant_rows(rows)
    G, b = rows[:, :4], rows[:, 4]
    assert 8 <= len(kept) <= len(rows) // 2, "most of the expanded rows are redundant"
    kept_set = set(int(k) for k in kept)
    for d in range(len(rows)):
        if d in kept_set:
            continue
        assert np.all(w[d] >= 0) and w[d].sum() > 1e-3
        assert all(int(k) in kept_set for k, wk in zip(idx[d], w[d]) if wk > 0)
        assert np.abs(G[d] - w[d] @ G[idx[d]]).max() <= 1e-11            # g_d is a non-negative combination of kept rows
        assert (w[d] * b[idx[d]]).sum() <= b[d] + 1e-9                     # ... whose bound is at least as tight
    # same set: points inside every kept row are inside every row
    rng = np.random.default_rng(0)
    g = np.array(env.goal, dtype=float)
    pts = g + rng.uniform(-1, 1, size=(200_000, 4)) * np.array([6.0, 2.0, 0.5, 3.0])
    in_kept = np.all(pts @ G[kept].T <= b[kept], axis=1)
    in_all = np.all(pts @ G.T <= b + 1e-9, axis=1)
    assert in_kept.sum() > 100
    assert np.array_equal(in_kept, in_all | in_kept) and not np.any(in_kept & ~in_all)
