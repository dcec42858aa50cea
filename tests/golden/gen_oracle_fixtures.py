"""Oracle-generated fixtures that are too slow to recompute inside the GPU suite.

Test infrastructure only.  Unlike ``gen_golden.py`` this does not need the reference tree: it runs the CPU oracle
(``oracle/carmpc_oracle.py``, exact QP solutions) and records its answers, so that the GPU test compares against
them without spending six minutes of host time on the GPU box.  ``tests/test_oracle_golden.py`` re-derives a subset
live so the fixture cannot drift from the oracle.

    python tests/golden/gen_oracle_fixtures.py

closed_loop_config4.npz - BASELINE config 4 in its own shape (RoadEnv, N = 20, 200 steps, initial states uniform in
x [0,10], y [-2.5,2.5], psi [-0.2,0.2], v [0,3], observer started at the true state, loop order of
examples/run_MPCOutputFB.py:29-41): 200 runs x 200 steps for output feedback and for state feedback; fail_step,
final state and every 10th trajectory step.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

RUNS, STEPS, STRIDE = 200, 200, 10


def config4_initial_states(runs: int = RUNS, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.uniform([0, -2.5, -0.2, 0], [10, 2.5, 0.2, 3], size=(runs, 4))


def run_oracle_loop(x_init, steps, output_feedback):
    from oracle import carmpc_oracle as orc
    Ab = np.load(os.path.join(HERE, "terminal_sets", "RoadEnv_30_1.5_0_0.npy"))
    oq = orc.CondensedQP("RoadEnv", 20, Ab)
    g = np.array([30, 1.5, 0, 0.0])

    def exact(x):
        u, obj, st, *_ = orc.qp_solve_exact(oq, x, g)
        return u[:, :2], st

    return orc.closed_loop(oq, x_init, g, steps, output_feedback, exact)


def main():
    x_init = config4_initial_states()
    out = {"x_init": x_init, "steps": STEPS, "stride": STRIDE}
    for tag, fb in (("ofb", True), ("sfb", False)):
        x, fail, traj = run_oracle_loop(x_init, STEPS, fb)
        out[f"final_{tag}"], out[f"fail_{tag}"], out[f"traj_{tag}"] = x, fail, traj[STRIDE - 1::STRIDE].copy()
        print(tag, "survivors", int((fail < 0).sum()), "of", RUNS)
    np.savez_compressed(os.path.join(HERE, "closed_loop_config4.npz"), **out)


if __name__ == "__main__":
    main()
