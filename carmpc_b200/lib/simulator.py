"""Nonlinear kinematic-bicycle plant.

Mirror of the reference's ``lib/simulator.py`` (``CarTrailerDimension`` :5-13, ``CarSimulator``
:16-118): forward-Euler step of [x, y, psi, v] under [a, delta_f], input assertion / clipping,
output y = C x (+ Cd d).  This class is the single-run host object; the batched Monte-Carlo closed
loop integrates the same equations on the GPU (``csrc/closed_loop.cu``).
"""
from typing import Union

import numpy as np


class CarTrailerDimension:
    l1 = 3.5
    l2 = 4
    l12 = 1
    car_length = 5
    car_width = 2
    trailer_length = 5.5
    trailer_width = 2
    triangle_length = 2


class CarSimulator:
    def __init__(self, dt: float = 0.01, clip: bool = False, C: np.ndarray = None,
                 Cd: np.ndarray = None) -> None:
        self.time = 0
        self.l1 = CarTrailerDimension.l1
        self.l2 = CarTrailerDimension.l2
        self.l12 = CarTrailerDimension.l12
        self.state = np.zeros(4)                       # [x, y, psi, v]
        self.dt = dt
        self.clip = clip
        self.C = np.eye(4) if C is None else C
        assert self.C.shape[1] == 4, \
            f"The length of the second dimension (currently {self.C.shape[1]}) of C should be equal to the state length (4)."
        self.output = self.C @ self.state
        self.Cd = Cd
        self.x_lower, self.x_upper = -np.inf, np.inf
        self.y_lower, self.y_upper = -np.inf, np.inf
        self.psi_lower, self.psi_upper = -np.pi, np.pi
        self.v_lower, self.v_upper = -10, 10
        # input limits, with the reference's 0.05 margin
        self.delta_lower, self.delta_upper = -np.pi / 4 - 0.05, np.pi / 4 + 0.05
        self.acc_lower, self.acc_upper = -2 - 0.05, 2 + 0.05

    def reset(self, state: Union[np.ndarray, list] = np.zeros(5)) -> None:
        assert np.array(state).shape == (4,), \
            f"The state should have shape (4,), but has shape{np.array(state).shape}"
        self.state = state

    def dynamics_continuous(self, state: np.ndarray, control_input: Union[np.ndarray, list]) -> np.ndarray:
        """One explicit-Euler step of the continuous bicycle model (reference :51-69)."""
        assert np.array(control_input).shape == (2,), \
            f"The input should have shape (2,), but has shape{np.array(control_input).shape}"
        _, _, psi, v = state
        a, delta_f = control_input
        rate = np.array([v * np.cos(psi), v * np.sin(psi), v / self.l1 * np.tan(delta_f), a])
        return rate * self.dt + state

    def get_log(self, control_input: Union[np.ndarray, list]) -> dict[str, np.ndarray]:
        a, delta_f = control_input
        return {'car': self.state, 'inputs': np.array([a, delta_f])}

    def check_input(self, control_input: Union[np.ndarray, list]) -> Union[np.ndarray, list]:
        """Assert (or clip, if ``clip``) the input against the actuator limits (reference :83-97).

        As in the reference, the clipped value is returned but ``step`` ignores the return value.
        """
        a, delta_f = control_input
        if self.clip:
            control_input = np.clip(control_input, [self.acc_lower, self.delta_lower],
                                    [self.acc_upper, self.delta_upper])
        else:
            assert np.all([a, -a] <= [self.acc_upper, -self.acc_lower]), \
                f"Acceleration should be between [{self.acc_lower, self.acc_upper}], but is {a}."
            assert np.all([delta_f, -delta_f] <= [self.delta_upper, -self.delta_lower]), \
                f"Steering angle should be between [{self.delta_lower:.3f}, {self.delta_upper:.3f}], but is {delta_f}."
        return control_input

    def step(self, control_input: Union[np.ndarray, list], d: float = None) -> dict[str, np.ndarray]:
        self.check_input(control_input)
        self.state = self.dynamics_continuous(self.state, control_input)
        self.output = self.C @ self.state
        if self.Cd is not None and d is not None:
            self.output += self.Cd * d
        self.time += self.dt
        return self.get_log(control_input)
