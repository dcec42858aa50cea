"""Nonlinear kinematic-bicycle plant (single run, host side).

Drop-in for the reference's ``lib/simulator.py`` (``CarTrailerDimension`` :5-13, ``CarSimulator`` :16-118):
same constructor, attributes and methods; ``step`` advances ``[x, y, psi, v]`` by one forward-Euler
step under ``[a, delta_f]``, after asserting (or, with ``clip=True``, clipping) the input against the
actuator limits, and refreshes ``output = C x (+ Cd d)``.  The batched Monte-Carlo closed loop
integrates the very same update on the GPU (``csrc/closed_loop.cu``, ``loop_advance_kernel``).
"""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np

ArrayLike = Union[np.ndarray, Sequence[float]]


class CarTrailerDimension:
    """Geometry in metres (only ``l1``, the wheelbase, enters the dynamics)."""
    l1 = 3.5
    l2 = 4
    l12 = 1
    car_length = 5
    car_width = 2
    trailer_length = 5.5
    trailer_width = 2
    triangle_length = 2


def bicycle_rate(state: ArrayLike, control_input: ArrayLike, wheelbase: float) -> np.ndarray:
    """Time derivative of ``[x, y, psi, v]``: ``[v cos psi, v sin psi, v / l1 * tan delta_f, a]``."""
    heading, speed = state[2], state[3]
    accel, steer = control_input
    return np.array([speed * np.cos(heading), speed * np.sin(heading), speed / wheelbase * np.tan(steer), accel])


class CarSimulator:
    #: slack the reference grants on top of the nominal actuator limits (2 m/s^2, pi/4 rad)
    INPUT_MARGIN = 0.05

    def __init__(self, dt: float = 0.01, clip: bool = False, C: np.ndarray = None, Cd: np.ndarray = None) -> None:
        dims = CarTrailerDimension
        self.l1, self.l2, self.l12 = dims.l1, dims.l2, dims.l12
        self.dt, self.clip, self.time = dt, clip, 0
        self.state = np.zeros(4)
        self.C = np.eye(4) if C is None else C
        assert self.C.shape[1] == 4, \
            f"C maps the 4 states to the outputs, so it needs 4 columns; it has {self.C.shape[1]}."
        self.Cd = Cd
        self.output = self.C @ self.state
        # state ranges (informational, as in the reference: never enforced)
        self.x_lower, self.x_upper = -np.inf, np.inf
        self.y_lower, self.y_upper = -np.inf, np.inf
        self.psi_lower, self.psi_upper = -np.pi, np.pi
        self.v_lower, self.v_upper = -10, 10
        # actuator limits
        self.acc_upper = 2 + self.INPUT_MARGIN
        self.acc_lower = -self.acc_upper
        self.delta_upper = np.pi / 4 + self.INPUT_MARGIN
        self.delta_lower = -self.delta_upper

    # ------------------------------------------------------------------------------------------
    def reset(self, state: ArrayLike = np.zeros(5)) -> None:
        assert np.array(state).shape == (4,), \
            f"A state is [x, y, psi, v]; got an array of shape {np.array(state).shape}."
        self.state = state

    def dynamics_continuous(self, state: np.ndarray, control_input: ArrayLike) -> np.ndarray:
        """The state one Euler step later: ``rate * dt + state`` (reference :51-69)."""
        assert np.array(control_input).shape == (2,), \
            f"An input is [a, delta_f]; got an array of shape {np.array(control_input).shape}."
        return bicycle_rate(state, control_input, self.l1) * self.dt + state

    def get_log(self, control_input: ArrayLike) -> dict:
        return {'car': self.state, 'inputs': np.array([control_input[0], control_input[1]])}

    def check_input(self, control_input: ArrayLike) -> ArrayLike:
        """Clip (``clip=True``) or assert the input against ``[acc_lower, acc_upper] x [delta_lower, delta_upper]``.
        Like the reference (:83-97) the possibly clipped value is returned, and ``step`` does not use it."""
        low = np.array([self.acc_lower, self.delta_lower])
        high = np.array([self.acc_upper, self.delta_upper])
        if self.clip:
            return np.clip(control_input, low, high)
        accel, steer = control_input
        assert low[0] <= accel <= high[0], \
            f"Acceleration {accel} is outside [{self.acc_lower}, {self.acc_upper}]."
        assert low[1] <= steer <= high[1], \
            f"Steering angle {steer} is outside [{self.delta_lower:.3f}, {self.delta_upper:.3f}]."
        return control_input

    def step(self, control_input: ArrayLike, d: float = None) -> dict:
        self.check_input(control_input)
        self.state = self.dynamics_continuous(self.state, control_input)
        self.output = self.C @ self.state
        if self.Cd is not None and d is not None:
            self.output += self.Cd * d
        self.time += self.dt
        return self.get_log(control_input)
