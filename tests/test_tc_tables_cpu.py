"""The tensor-core form of the ADMM (csrc/qp_admm_tc.cu) against the FFMA form, on the host: the swizzled TF32 hi / lo chunk
images and the shifted iteration (w^ = w - h, constant columns) must reproduce the FFMA kernel's iterates (both re-run in
numpy from the tables the library exports; no GPU)."""
import numpy as np
import pytest

from conftest import make_env, make_controller
from admm_emulation import KernelTables, emulate
from tc_emulation import TcTables, emulate_tc


@pytest.mark.parametrize("env_name,goal,N", [("RoadOneCarEnv", [29.9, 1.5, 0, 0], 10), ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20),
                                             ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 40), ("RoadMultipleCarsEnv", [30, 1.5, 0, 0], 20),
                                             ("RoadEnv", [30, 1.5, 0, 0], 20)])
def test_tensor_form_reproduces_ffma_iterates(env_name, goal, N):
    from carmpc_b200.batch import BatchQP
    c = make_controller(make_env(env_name, goal), N)
    bq = BatchQP.from_controller(c)
    T, C = KernelTables(bq), TcTables(bq)
    assert C.ok, "horizons up to 40 have a tensor-core form"
    # geometry the kernel relies on: MMA N multiples of 16, chunks cover K, tensor memory budget, shared memory budget
    assert all(int(v) % 16 == 0 and 16 <= int(v) <= 256 for v in C.ncols)
    assert 2 * C.mp + C.np_ <= 512 and C.smem <= 227 * 1024
    assert list(C.ksteps) == [(C.np_ + 16 + C.mp) // 8, (C.np_ + 16) // 8, C.mp // 8]
    goal_v = np.array(c.goal, float)
    rng = np.random.default_rng(5)
    lo = np.array([5.0, -3.0, -np.pi / 8, -1.0]); hi = np.array([30.0, 3.0, np.pi / 8, 5.0])
    x0 = lo + rng.uniform(size=(96, 4)) * (hi - lo)
    for iters in (1, 25):
        _, sign, w = emulate(T, x0, goal_v, iters, return_state=True)
        w2, _, _, sign2 = emulate_tc(C, x0, goal_v, iters)
        assert np.abs(w - w2).max() <= 2e-5 * max(1.0, np.abs(w).max())
        assert (sign == sign2).all(1).mean() >= 0.97


def test_horizon_80_splits_into_its_independent_chains():
    """346 general rows do not fit tensor memory (state + accumulators exceed 512 columns), but K is block diagonal over the
    acceleration and steering chains of RoadOneCarEnv: each chain is a horizon-40-sized part.  RoadMultipleCarsEnv couples
    the chains (its terminal / car rows mix x and y): no tensor-core form at horizon 80."""
    from carmpc_b200.batch import BatchQP
    bq = BatchQP.from_controller(make_controller(make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0]), 80))
    part0 = TcTables(bq)                               # the library exports the first part
    assert part0.ok and part0.np_ == 80 and part0.mp <= 192 and 2 * part0.mp + part0.np_ <= 512
    live = part0.row_id[part0.row_id >= 0]
    assert 80 < len(live) < bq.pq.m                    # a strict subset of the rows: one chain
    bq2 = BatchQP.from_controller(make_controller(make_env("RoadMultipleCarsEnv", [30, 1.5, 0, 0]), 80))
    assert not TcTables(bq2).ok
