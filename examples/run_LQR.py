#!/usr/bin/env python
"""Counterpart of the reference's examples/run_LQR.py (headless): the unconstrained LQR law of the controller
(``use_LQR=True``, lib/mpc.py:255-268) on the nonlinear bicycle with input clipping.

    python examples/run_LQR.py [--start 25 1.0 0 2] [--seconds 20]
"""
import argparse

import numpy as np

import _common
from carmpc_b200.lib.configuration import DT_CONTROL, DT_SIMULATION, STEPS_UPDATE, LINEARIZE_STATE, LINEARIZE_INPUT, N
from carmpc_b200.lib.mpc import MPC
from carmpc_b200.lib.simulator import CarSimulator


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--start", type=float, nargs=4, default=[25, 1.0, 0, 2])
    ap.add_argument("--seconds", type=float, default=20.0)
    args = ap.parse_args(argv)
    env = _common.make_env("RoadEnv")
    controller = MPC(dt=DT_CONTROL, N=N, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, env=env, use_LQR=True)
    controller.set_goal(env.goal)
    plant = CarSimulator(dt=DT_SIMULATION, clip=True)
    plant.reset(np.array(args.start, dtype=float))
    u = np.zeros(2)
    for i in range(int(args.seconds / DT_SIMULATION) + 1):
        plant.step(u)
        if i % STEPS_UPDATE == 0:
            u = controller.step(plant.state)
        if np.all(np.abs(plant.state - controller.goal) <= 1e-1):
            break
    reached = bool(np.all(np.abs(plant.state - controller.goal) <= 1e-1))
    print(f"LQR: t = {plant.time:.1f} s, final state {np.round(plant.state, 3)}, reached = {reached}")
    return plant.state, reached


if __name__ == "__main__":
    main()
