#!/usr/bin/env python
"""What the batched GPU path adds to the reference: the region of attraction of the state-feedback MPC as a sampled map
(one condensed QP per grid point, ``carmpc_qp_map_host``), its hull polytope in the terminal_sets/*.npy layout, and a
Monte-Carlo closed loop of the output-feedback controller against the nonlinear bicycle (``carmpc_closed_loop``).

    python examples/roa_map.py [--points 60 60 6 6] [--runs 20000] [--steps 200]
"""
import argparse
import time

import numpy as np

import _common
from carmpc_b200 import roa
from carmpc_b200.batch import BatchQP
from carmpc_b200.lib.configuration import DT_CONTROL, LINEARIZE_STATE, LINEARIZE_INPUT, N
from carmpc_b200.lib.mpc import MPCStateFB, MPCOutputFB


def main(argv=None):
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, nargs=4, default=[60, 60, 6, 6])
    ap.add_argument("--runs", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args(argv)
    env = _common.make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0])
    controller = MPCStateFB(dt=DT_CONTROL, N=N, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, env=env)
    nx, ny, npsi, nv = args.points
    axes = [np.linspace(5.0, 30.0, nx), np.linspace(-3.0, 3.0, ny), np.linspace(-np.pi / 8, np.pi / 8, npsi), np.linspace(-1.0, 5.0, nv)]
    t0 = time.perf_counter()
    A, b, flags = roa.region_of_attraction(controller, axes)
    dt = time.perf_counter() - t0
    print(f"region of attraction: {int(flags.sum())} of {flags.size} grid states feasible, hull polytope with {len(b)} facets "
          f"({dt:.2f} s including the hull)")

    ofb_env = _common.make_env("RoadEnv")
    ofb = MPCOutputFB(dt=DT_CONTROL, N=N, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, init_state=[0, 0, 0, 2], env=ofb_env)
    qp = BatchQP.from_controller(ofb)
    g = torch.Generator(device="cpu").manual_seed(0)
    lo = torch.tensor([0.0, -2.5, -0.2, 0.0], dtype=torch.float64)
    hi = torch.tensor([10.0, 2.5, 0.2, 3.0], dtype=torch.float64)
    x_init = (lo[:, None] + (hi - lo)[:, None] * torch.rand((4, args.runs), generator=g, dtype=torch.float64)).cuda().contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = qp.closed_loop(x_init, args.steps, ofb.A, ofb.B, C=ofb.C, L=ofb.L)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = out["fail_step"] < 0
    goal = torch.tensor(np.asarray(ofb.goal, dtype=float), device="cuda")
    reached = ((out["final"] - goal[:, None]).abs() <= 0.1).all(0) & ok
    print(f"closed loop: {args.runs} runs x {args.steps} steps in {dt:.3f} s, {int(ok.sum())} never infeasible, "
          f"{int(reached.sum())} at the goal")
    return A, b, flags, out


if __name__ == "__main__":
    main()
