"""QP half of ``__graft_entry__.smoke()``: a small batch of horizon-20 QPs on cuda:0, checked against the oracle."""
import os

import numpy as np


def run(root: str) -> None:
    from oracle import carmpc_oracle as orc
    from .batch import BatchQP
    from .lib.environments import RoadOneCarEnv
    from .lib.mpc import MPCStateFB
    from .lib.configuration import DT_CONTROL, LINEARIZE_STATE, LINEARIZE_INPUT
    from .lib import terminal_set as ts
    env = RoadOneCarEnv()
    env.set_goal([29.9, 1.5, 0, 0])
    ts.TERMINAL_SET_DIR = os.path.join(root, "terminal_sets")
    controller = MPCStateFB(dt=DT_CONTROL, N=20, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, env=env)
    bq = BatchQP.from_controller(controller)
    Ab = np.load(os.path.join(root, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    oq = orc.CondensedQP("RoadOneCarEnv", 20, Ab)
    goal = np.array(env.goal, dtype=float)
    rng = np.random.default_rng(0)
    x0 = goal + rng.uniform(-1, 1, size=(64, 4)) * np.array([12.0, 1.4, 0.25, 2.5])
    res = bq.solve_host(x0)
    ue, obje, ste, polished, slack = orc.qp_solve_exact(oq, x0, goal)
    band = np.abs(slack) <= 1e-6
    assert np.array_equal(np.where(res.status == 0, 0, 1)[~band], ste[~band]), "feasibility flags differ from the oracle"
    ok = (ste == 0) & ~band
    du = np.abs(res.u0[ok] - ue[ok, :2]).max()
    rel = (np.abs(res.objective[ok] - obje[ok]) / np.maximum(1, np.abs(obje[ok]))).max()
    assert du <= 1e-4 and rel <= 1e-5, (du, rel)
    print(f"smoke: QP ok ({int(ok.sum())} feasible of {len(x0)}, max |du0| {du:.2e}, max rel dobj {rel:.2e}, "
          f"mean ADMM iterations {res.iters.mean():.1f})")
