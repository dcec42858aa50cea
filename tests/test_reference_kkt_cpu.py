"""The oracle's exact QP answers are optimal for the REFERENCE's problem (CPU suite).

``kkt_check.ReferenceQP`` assembles the QP from matrices the unmodified reference controller produced
(tests/golden/qp_*.npz, lib/mpc.py:318-332) and checks optimality conditions directly; it shares no code with the
oracle's matrix builder or solvers.  This pins the oracle's problem definition, objective and feasibility flags to the
reference for every BASELINE configuration (N = 10/20/40/80, three environments); the GPU suite applies the same
checker to the CUDA results (tests/test_qp_gpu.py::test_gpu_results_are_kkt_points_of_the_reference_qp).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from kkt_check import ReferenceQP, assert_results_satisfy_reference_qp

CASES = [("RoadOneCarEnv", 10, "RoadOneCarEnv_29.9_1.5_0_0.npy", 60), ("RoadOneCarEnv", 20, "RoadOneCarEnv_29.9_1.5_0_0.npy", 60),
         ("RoadOneCarEnv", 40, "RoadOneCarEnv_29.9_1.5_0_0.npy", 24), ("RoadOneCarEnv", 80, "RoadOneCarEnv_29.9_1.5_0_0.npy", 10),
         ("RoadEnv", 20, "RoadEnv_30_1.5_0_0.npy", 40), ("RoadMultipleCarsEnv", 20, "RoadMultipleCarsEnv_30_1.5_0_0.npy", 40)]


@pytest.mark.parametrize("env_name,N,fixture,n_states", CASES)
def test_oracle_solutions_are_kkt_points_of_the_reference_qp(env_name, N, fixture, n_states):
    from oracle import carmpc_oracle as orc
    ref = ReferenceQP(env_name, N)
    oq = orc.CondensedQP(env_name, N, np.load(os.path.join(GOLDEN, "terminal_sets", fixture)))
    # the oracle's own assembly equals the reference's, row for row
    np.testing.assert_allclose(oq.H, ref.H, rtol=0, atol=1e-9 * np.abs(ref.H).max())
    np.testing.assert_allclose(oq.G, ref.G, rtol=0, atol=1e-10 * np.abs(ref.G).max())
    rng = np.random.default_rng(N)
    x0 = ref.goal + rng.uniform(-1, 1, size=(n_states, 4)) * np.array([12.0, 1.4, 0.25, 2.5])
    np.testing.assert_allclose(oq.rhs(x0), ref.rhs(x0), rtol=0, atol=1e-9)
    u, obj, status, polished, slack = orc.qp_solve_exact(oq, x0, ref.goal)
    assert polished[status == 0].all()
    info = assert_results_satisfy_reference_qp(ref, x0, u, status, objective=np.where(status == 0, obj, 0.0), n_lp=30)
    assert info["solved"] >= 2 and info["infeasible"] >= 1, info
    # both flag sources agree with the reference LP on every state outside the band
    lp = ref.feasibility_slack(x0)
    band = np.abs(lp) <= 1e-6
    np.testing.assert_array_equal((lp >= 0)[~band], (status == 0)[~band])
    np.testing.assert_allclose(lp[~band & (lp < 1)], slack[~band & (lp < 1)], atol=1e-7)


def test_checker_rejects_wrong_answers():
    """The checker is not vacuous: a perturbed optimum, a dropped constraint and a flipped flag all fail."""
    from oracle import carmpc_oracle as orc
    ref = ReferenceQP("RoadOneCarEnv", 20)
    oq = orc.CondensedQP("RoadOneCarEnv", 20, np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy")))
    x0 = np.array([[12.0, 0.4, 0.05, 2.0], [20.0, 1.0, 0.0, 1.0]])
    u, obj, status, *_ = orc.qp_solve_exact(oq, x0, ref.goal)
    assert (status == 0).all()
    assert_results_satisfy_reference_qp(ref, x0, u, status, objective=obj)
    bumped = u.copy()
    bumped[0, 5] += 1e-4
    with pytest.raises(AssertionError):
        assert_results_satisfy_reference_qp(ref, x0, bumped, status)
    with pytest.raises(AssertionError, match="primal"):
        assert_results_satisfy_reference_qp(ref, x0, u + 0.5, status)
    with pytest.raises(AssertionError, match="strictly feasible"):
        assert_results_satisfy_reference_qp(ref, x0, u, np.array([0, 1]))
