#!/usr/bin/env python
"""Counterpart of the reference's examples/run_MPCStateFB.py (headless): drive the nonlinear bicycle to the goal with the
state-feedback MPC, one QP per control step on the GPU through the drop-in ``MPCStateFB.step``.

    python examples/run_MPCStateFB.py [--env RoadMultipleCarsEnv] [--start 5 -1.5 0.1 0] [--seconds 30]
"""
import argparse

import numpy as np

import _common
from carmpc_b200.lib.configuration import DT_CONTROL, DT_SIMULATION, STEPS_UPDATE, LINEARIZE_STATE, LINEARIZE_INPUT, N
from carmpc_b200.lib.mpc import MPCStateFB, OutsideTheRegionOfAttractionError
from carmpc_b200.lib.simulator import CarSimulator


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="RoadMultipleCarsEnv")
    ap.add_argument("--start", type=float, nargs=4, default=[5, -1.5, 0.1, 0])
    ap.add_argument("--seconds", type=float, default=30.0)
    args = ap.parse_args(argv)
    env = _common.make_env(args.env)
    controller = MPCStateFB(dt=DT_CONTROL, N=N, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, env=env)
    controller.set_goal(env.goal)
    plant = CarSimulator(dt=DT_SIMULATION)
    plant.reset(np.array(args.start, dtype=float))
    u = np.zeros(2)
    states, inputs, costs = [], [], []
    for i in range(int(args.seconds / DT_SIMULATION) + 1):
        log = plant.step(u)
        states.append(log["car"])
        inputs.append(log["inputs"])
        if i % STEPS_UPDATE == 0:
            try:
                u = controller.step(plant.state)
            except OutsideTheRegionOfAttractionError:
                print(f"t = {plant.time:.1f} s: state {np.round(plant.state, 3)} is outside the region of attraction")
                break
            costs.append(controller.stage_cost + controller.terminal_cost)
        if np.all(np.abs(plant.state - controller.goal) <= 1e-1):
            break
    states = np.array(states)
    reached = bool(np.all(np.abs(plant.state - controller.goal) <= 1e-1))
    print(f"{args.env}: {len(states)} plant steps, final state {np.round(plant.state, 3)}, goal {np.round(controller.goal, 3)}, "
          f"reached = {reached}, max |a| {np.abs(np.array(inputs)[:, 0]).max():.2f}, max |delta| {np.abs(np.array(inputs)[:, 1]).max():.3f}")
    return states, np.array(inputs), np.array(costs), reached


if __name__ == "__main__":
    main()
