#!/usr/bin/env python
"""Counterpart of the reference's examples/run_MPCOutputFB.py (headless): output feedback - the controller sees (x, y, v)
only, a Luenberger observer reconstructs the heading (lib/mpc.py:439-448).

    python examples/run_MPCOutputFB.py [--env RoadMultipleCarsEnv] [--start 5 -1.5 0 0] [--seconds 30]
"""
import argparse

import numpy as np

import _common
from carmpc_b200.lib.configuration import DT_CONTROL, DT_SIMULATION, STEPS_UPDATE, LINEARIZE_STATE, LINEARIZE_INPUT, N
from carmpc_b200.lib.mpc import MPCOutputFB, OutsideTheRegionOfAttractionError
from carmpc_b200.lib.simulator import CarSimulator


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="RoadMultipleCarsEnv")
    ap.add_argument("--start", type=float, nargs=4, default=[5, -1.5, 0, 0])
    ap.add_argument("--seconds", type=float, default=30.0)
    args = ap.parse_args(argv)
    env = _common.make_env(args.env)
    start = np.array(args.start, dtype=float)
    controller = MPCOutputFB(dt=DT_CONTROL, N=N, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, init_state=start, env=env)
    controller.set_goal(env.goal)
    plant = CarSimulator(dt=DT_SIMULATION, C=controller.C)
    plant.reset(start)
    u = np.zeros(2)
    err = []
    for i in range(int(args.seconds / DT_SIMULATION) + 1):
        plant.step(u)
        if i % STEPS_UPDATE == 0:
            try:
                u = controller.step(plant.output)
            except OutsideTheRegionOfAttractionError:
                print(f"t = {plant.time:.1f} s: estimate {np.round(controller.x_estimate, 3)} is outside the region of attraction")
                break
            err.append(np.abs(controller.x_estimate - plant.state).max())
        if np.all(np.abs(plant.state - controller.goal) <= 1e-1):
            break
    reached = bool(np.all(np.abs(plant.state - controller.goal) <= 1e-1))
    print(f"{args.env}: t = {plant.time:.1f} s, final state {np.round(plant.state, 3)}, reached = {reached}, "
          f"largest observer error {max(err, default=float('nan')):.3e}")
    return plant.state, reached, np.array(err)


if __name__ == "__main__":
    main()
