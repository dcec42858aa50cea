"""Module-level constants of the controller and simulator.

Mirror of the reference's ``lib/configuration.py:1-21`` (same names and values); star-imported by
``mpc`` and ``terminal_set`` exactly as in the reference (``lib/mpc.py:14``).
"""
import numpy as np

DT_CONTROL = 0.2                                  # MPC update period [s] (zero-order hold on u)
DT_SIMULATION = DT_CONTROL                        # plant integration step [s]
STEPS_UPDATE = int(DT_CONTROL / DT_SIMULATION)    # plant steps per controller step

LINEARIZE_STATE = [0, 0, 0, 3]                    # [x, y, psi, v] the bicycle is linearised around
LINEARIZE_INPUT = [0, 0]                          # [a, delta_f]

N = 20                                            # prediction horizon

STAGE_COST_Q = np.diag([5, 5, 10, 10])
STAGE_COST_R = np.diag([10, 100])
