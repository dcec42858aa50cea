"""The exact Farkas certificates of csrc/qp_polish.cu (farkas_kernel), restated in numpy on the float32 ADMM state the
kernel emulation produces.  What the GPU path relies on:

* soundness - a certificate is accepted only if S(x0) < 0, and S < 0 implies infeasibility whatever y is (A'y = 0 holds by
  construction because the box multipliers absorb the residual), so no feasible state may ever be "proven" infeasible;
* the support value is affine in the state, S(x0) = c0 - cx.x0, which is what lets a follower reuse its anchor's y;
* usefulness - the dual iterate of an infeasible state is a valid certificate for most infeasible states and neighbours.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_env, make_controller
from admm_emulation import KernelTables, emulate


@pytest.fixture(scope="module")
def problem():
    from carmpc_b200.batch import BatchQP
    from oracle import carmpc_oracle as orc
    c = make_controller(make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0]), 20)
    bq = BatchQP.from_controller(c)                      # host-side setup only (no CUDA device needed)
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    oq = orc.CondensedQP("RoadOneCarEnv", 20, Ab)
    goal = np.array(c.goal, dtype=float)
    xs, ys = np.linspace(5.0, 30.0, 12), np.linspace(-3.0, 3.0, 25)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    x0 = np.stack([X.ravel(), Y.ravel(), np.full(X.size, 0.2), np.full(X.size, 3.0)], axis=1)
    feasible, slack = orc.qp_feasible_lp(oq, x0)
    T = KernelTables(bq)
    _, _, w = emulate(T, x0, goal, 30, return_state=True)
    return bq.pq, bq.setup(1), x0, feasible, slack, w, (len(xs), len(ys))


def _certificate(pq, Eg, w, x0):
    """y = E (w - clip(w)) at the sample's own bounds, box multipliers y_b = -G'y; returns (cx (B, 4), c0, Ax, A0)."""
    hi = pq.hi[None, :] - x0 @ pq.Gx.T
    lo = pq.lo[None, :] - x0 @ pq.Gx.T
    ws = w.astype(float)
    hs, ls = Eg[None, :] * hi, Eg[None, :] * lo
    y = Eg[None, :] * np.where(ws > hs, ws - hs, np.where(ws < ls, ws - ls, 0.0))
    with np.errstate(invalid="ignore"):
        bound = np.where(y > 0, pq.hi[None, :], np.where(y < 0, pq.lo[None, :], 0.0))
        t = np.where(y != 0, bound * y, 0.0)
    yb = -(y @ pq.G)
    tb = np.where(yb > 0, pq.ub[None, :] * yb, np.where(yb < 0, pq.lb[None, :] * yb, 0.0))
    cx = y @ pq.Gx
    c0 = t.sum(1) + tb.sum(1)
    A0 = np.abs(t).sum(1) + np.abs(tb).sum(1)
    Ax = np.abs(y) @ np.abs(pq.Gx)
    return y, yb, cx, c0, Ax, A0


def _support_direct(pq, y, yb, x0):
    hi = pq.hi[None, :] - x0 @ pq.Gx.T
    lo = pq.lo[None, :] - x0 @ pq.Gx.T
    with np.errstate(invalid="ignore"):
        tg = np.where(y > 0, hi * y, np.where(y < 0, lo * y, 0.0))
    tb = np.where(yb > 0, pq.ub[None, :] * yb, np.where(yb < 0, pq.lb[None, :] * yb, 0.0))
    return tg.sum(1) + tb.sum(1)


def test_certificates_are_sound_affine_and_useful(problem):
    pq, Eg, x0, feasible, slack, w, shape = problem
    assert 20 < feasible.sum() < len(x0) - 20
    y, yb, cx, c0, Ax, A0 = _certificate(pq, Eg, w, x0)
    # A'y = 0 by construction (general rows + identity box rows)
    assert np.abs(y @ pq.G + yb).max() == 0.0
    # affine form == direct evaluation of the support function, at the own state and at a shifted one
    S = c0 - np.einsum("bi,bi->b", cx, x0)
    np.testing.assert_allclose(S, _support_direct(pq, y, yb, x0), rtol=0, atol=1e-9 * (1 + A0.max()))
    shifted = x0 + np.array([0.4, -0.2, 0.01, 0.1])
    S_shift = c0 - np.einsum("bi,bi->b", cx, shifted)
    np.testing.assert_allclose(S_shift, _support_direct(pq, y, yb, shifted), rtol=0, atol=1e-9 * (1 + A0.max()))
    margin = 1e-9 * (A0 + np.einsum("bi,bi->b", Ax, np.abs(x0)))
    valid = np.isfinite(S) & (S < -margin)
    # soundness: never for a feasible state (states within 1e-6 of the boundary excluded, as everywhere)
    assert not (valid & feasible & (np.abs(slack) > 1e-6)).any()
    # usefulness: most infeasible states are certified by their own dual iterate after 30 iterations
    infeasible = ~feasible & (np.abs(slack) > 1e-6)
    assert valid[infeasible].mean() >= 0.7, valid[infeasible].mean()


def test_an_anchor_certificate_is_sound_and_useful_for_its_neighbours(problem):
    pq, Eg, x0, feasible, slack, w, shape = problem
    y, yb, cx, c0, Ax, A0 = _certificate(pq, Eg, w, x0)
    nx, ny = shape
    idx = np.arange(len(x0)).reshape(nx, ny)
    proven = tried = 0
    for d in (1, 2, 4):
        anchor, follower = idx[:, :-d].ravel(), idx[:, d:].ravel()          # the follower is d grid steps along y
        xf = x0[follower]
        S = c0[anchor] - np.einsum("bi,bi->b", cx[anchor], xf)
        margin = 1e-9 * (A0[anchor] + np.einsum("bi,bi->b", Ax[anchor], np.abs(xf)))
        ok = np.isfinite(S) & (S < -margin)
        clear = np.abs(slack[follower]) > 1e-6
        assert not (ok & feasible[follower] & clear).any(), "an anchor's certificate 'proved' a feasible follower infeasible"
        both = ~feasible[anchor] & ~feasible[follower] & clear
        proven += int(ok[both].sum())
        tried += int(both.sum())
    assert tried > 50 and proven / tried >= 0.6, (proven, tried)
