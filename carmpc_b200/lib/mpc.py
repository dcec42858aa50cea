"""MPC controllers of the linearised kinematic bicycle, solved on the GPU.

Same classes, constructor signatures, attributes and ``step`` semantics as the reference's
``lib/mpc.py``: ``MPC`` (:21-276), ``MPCStateFB`` (:279-349), ``MPCOutputFB`` (:352-492),
``MPCOutputFBWithDisturbance`` (:495-667) and ``OutsideTheRegionOfAttractionError`` (:17).

What differs is where the QP goes.  The reference rebuilds a cvxpy problem and calls
``problem.solve()`` (-> OSQP) on every step (:334-335, :477-478).  Here the condensed problem is
assembled once per controller (``carmpc_b200.condensed``), handed to the CUDA batched-ADMM solver
through the C ABI (``carmpc_qp_create``) and ``step`` is a batch-of-one call of the same kernel
that evaluates 10^6 states at a time (``carmpc_b200.batch.BatchQP``).  There is no CPU solver in
this package: without the CUDA library ``step`` raises.
"""
from __future__ import annotations

import os
from typing import Union

import numpy as np
import numpy.linalg
from scipy.linalg import solve_discrete_are as dare

from .simulator import CarTrailerDimension
from .matrix_gen import predmod, costgen
from .environments import BaseEnv
from .configuration import *          # noqa: F401,F403

#: extra directories searched for ``<env>_<goal>.npy`` after the reference's '../terminal_sets/'
TERMINAL_SET_SEARCH_PATH = [os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))), 'terminal_sets')]


class OutsideTheRegionOfAttractionError(Exception):
    """The QP is infeasible: the initial state lies outside the region of attraction."""


class MPC:
    def __init__(self,
                 dt: float,
                 N: int,
                 lin_state: Union[np.ndarray, list],
                 lin_input: Union[np.ndarray, list],
                 terminal_constraint: bool = True,
                 input_constraint: bool = True,
                 state_constraint: bool = True,
                 env: BaseEnv = None,
                 use_LQR: bool = False,
                 ) -> None:
        self.dt = dt
        self.N = N
        self.lin_state = np.array(lin_state) if isinstance(lin_state, list) else lin_state
        self.lin_input = np.array(lin_input) if isinstance(lin_input, list) else lin_input
        self.goal = np.array(env.goal)
        self.x_horizon, self.u_horizon = None, None
        self.cost = 0
        self.env = env
        self.use_LQR = use_LQR

        A_lin, B_lin = self.linearized_model(self.lin_state, self.lin_input)
        self.A, self.B = self.discretized_model(A_lin, B_lin, dt)

        self.state_const_A, self.state_const_b = [], []
        self.nx, self.nu = 4, 2
        self.Q = STAGE_COST_Q
        self.R = STAGE_COST_R

        self.terminal_constraint_bool = terminal_constraint
        self.input_constraint_bool = input_constraint
        self.state_constraint_bool = state_constraint

        try:
            self.P = dare(self.A, self.B, self.Q, self.R)
        except numpy.linalg.LinAlgError:
            print("\nNOT POSSIBLE TO SOLVE DARE\n")
            self.P = np.zeros((4, 4))

        self.T, self.S = predmod(self.A, self.B, self.N)
        self.H, self.h, _ = costgen(self.Q, self.R, self.P, self.T, self.S, self.nx)

        # unconstrained optimal law u = K (x - goal)
        self.K = -np.linalg.inv(self.R + self.B.T @ self.P @ self.B) @ self.B.T @ self.P @ self.A

        # inputs [acceleration, steering angle]
        self.input_upper = np.array([2, np.pi / 8])
        self.input_lower = -self.input_upper

        # stay close to the linearisation point: |psi| <= pi/8, -1 <= v <= 5
        self.state_const_A += [[0, 0, 1, 0], [0, 0, -1, 0]]
        self.state_const_b += [np.pi / 8, np.pi / 8]
        self.state_const_A += [[0, 0, 0, 1], [0, 0, 0, -1]]
        self.state_const_b += [5, 1]

        self._qp = None                 # lazily created GPU solver (carmpc_b200.batch.BatchQP)

        if terminal_constraint:
            self._load_terminal_set(env)

    # ------------------------------------------------------------------ terminal set ----------
    def _load_terminal_set(self, env: BaseEnv) -> None:
        """Look the H-rep up by environment name and goal (reference :96-117).

        File layout: float64 (rows, 5) = [A | b], A x <= b in absolute coordinates.  When no file
        exists the goal itself (x, y, psi) becomes the terminal constraint, as in the reference.
        """
        filename = env.name + ''.join('_' + str(s) for s in env.goal) + '.npy'
        from . import terminal_set as _ts
        for directory in [_ts.TERMINAL_SET_DIR] + TERMINAL_SET_SEARCH_PATH:
            file = os.path.join(directory, filename)
            if os.path.isfile(file):
                terminal_set = np.load(file)
                A, b = terminal_set[..., :4], terminal_set[..., 4]
                A_hat = np.zeros((len(terminal_set), (self.N + 1) * self.nx))
                A_hat[:, -4:] = A
                self.term_set_A, self.term_set_b = A_hat, b
                print(f"Using {file} for the terminal set.")
                return
        print(f"No terminal set found for the specific combination of environment{env.name} and goal{env.goal}.",
              f"The goal state will be used as terminal constraint, but this might be too strict.")
        sel = np.diag([1, 1, 1, 0])
        A = np.zeros((2 * self.nx, (self.N + 1) * self.nx))
        A[:self.nx, self.N * self.nx:] = sel
        A[self.nx:, self.N * self.nx:] = -sel
        self.term_set_A, self.term_set_b = A, np.hstack((self.goal, -self.goal))

    def check_constraints(self):
        assert len(self.state_const_A) == len(self.state_const_b), \
            f"The constraints array (len(A) = {len(self.state_const_A)} and len(B) = {len(self.state_const_b)})" + \
            "do not have the same length. This will produced unpredictable behaviour."

    # ------------------------------------------------------------------ model -----------------
    @staticmethod
    def linearized_model(lin_state: np.ndarray, lin_input: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """Jacobians of [v cos psi, v sin psi, v/l1 tan delta, a] at (lin_state, lin_input)."""
        assert lin_state.shape == (4,), f"The state should have shape [4], but has {lin_state.shape}."
        assert lin_input.shape == (2,), f"The input should have shape [2], but has {lin_input.shape}."
        _, _, psi_e, v_e = lin_state
        _, delta_e = lin_input
        l1 = CarTrailerDimension.l1
        A_lin = np.zeros((4, 4))
        A_lin[0, 2], A_lin[0, 3] = -v_e * np.sin(psi_e), np.cos(psi_e)
        A_lin[1, 2], A_lin[1, 3] = v_e * np.cos(psi_e), np.sin(psi_e)
        A_lin[2, 3] = np.tan(delta_e) / l1
        B_lin = np.zeros((4, 2))
        B_lin[2, 1] = v_e / l1 * 1 / np.cos(delta_e) ** 2
        B_lin[3, 0] = 1
        return A_lin, B_lin

    @staticmethod
    def discretized_model(A: np.ndarray, B: np.ndarray, dt: float) -> tuple[np.ndarray, np.ndarray]:
        """Forward Euler: A_d = I + dt A, B_d = dt B."""
        assert len(A) == len(B), "Matrices A and B should have the same height."
        return dt * A + np.eye(len(A)), dt * B

    def set_goal(self, goal: Union[np.ndarray, list]) -> None:
        assert np.array(goal).shape == (4,)
        assert np.all(goal == self.A @ goal), \
            "Goal is not an equilibrium point. At least not for the linearized model. E.g. psi != 0 is not an equilibrium point."
        self.goal = np.array(goal) if isinstance(goal, list) else goal
        self._qp = None

    # ------------------------------------------------------------------ constraint stacks -----
    def terminal_constraint(self) -> tuple[np.ndarray, np.ndarray]:
        """A x_ <= b with A non-zero only on x(N)."""
        return self.term_set_A, self.term_set_b

    def input_constraint(self) -> tuple[np.ndarray, np.ndarray]:
        """[I; -I] u_ <= [ub x N; -lb x N]."""
        assert self.nu == len(self.input_upper), \
            f"Number of inputs ({self.nu}) is different from number of upper inputs boundaries ({len(self.input_upper)})"
        assert self.nu == len(self.input_lower), \
            f"Number of inputs ({self.nu}) is different from number of upper inputs boundaries ({len(self.input_lower)})"
        eye = np.eye(self.N * self.nu)
        return np.vstack((eye, -eye)), np.hstack((np.tile(self.input_upper, self.N),
                                                  np.tile(-self.input_lower, self.N)))

    def state_constraint(self) -> tuple[np.ndarray, np.ndarray]:
        """Row type by row type (built-in rows first, then the environment's), N rows each, row i
        acting on x(i+1); x(0) is never constrained (reference :229-253)."""
        rows = list(zip(self.state_const_A, self.state_const_b))
        if self.env is not None:
            rows += list(zip(self.env.constraints_A, self.env.constraints_b))
        shift = np.eye(self.N, self.N + 1, k=1)          # picks block i+1 for row i
        A = np.vstack([np.kron(shift, np.asarray(a, dtype=float)[None, :]) for a, _ in rows])
        b = np.hstack([np.full(self.N, float(bb)) for _, bb in rows])
        return A, b

    # ------------------------------------------------------------------ control laws ----------
    def LQR(self, x0) -> np.ndarray:
        return self.K @ (x0 - self.goal)

    def step(self, x0) -> np.ndarray:
        if self.use_LQR:
            u0 = np.clip(self.LQR(x0), self.input_lower, self.input_upper)
            self.stage_cost = (x0 - self.goal) @ self.Q @ (x0 - self.goal) + u0 @ self.R @ u0
            return u0
        raise NotImplementedError

    # ------------------------------------------------------------------ GPU hook --------------
    def batch_solver(self):
        """The GPU solver bound to this controller's condensed QP (created on first use)."""
        if self._qp is None:
            from ..batch import BatchQP
            self._qp = BatchQP.from_controller(self)
        return self._qp

    def _solve_one(self, x0: np.ndarray, x_ref: np.ndarray):
        """One QP through the batched kernel.  Returns (u_ (2N,), cost); raises when infeasible."""
        res = self.batch_solver().solve_host(np.asarray(x0, dtype=float)[None, :],
                                             x_ref=np.asarray(x_ref, dtype=float), want_u_full=True)
        if res.status[0] == 1:
            raise OutsideTheRegionOfAttractionError
        if res.status[0] != 0:
            raise RuntimeError(f"QP solver stopped with status {int(res.status[0])} after "
                               f"{int(res.iters[0])} iterations")
        return res.u_full[0], float(res.objective[0])


class MPCStateFB(MPC):
    def __init__(self, dt, N, lin_state, lin_input, terminal_constraint: bool = True,
                 input_constraint: bool = True, state_constraint: bool = True,
                 env: BaseEnv = None) -> None:
        super().__init__(dt, N, lin_state, lin_input, terminal_constraint, input_constraint,
                         state_constraint, env)
        self.check_constraints()

    def step(self, x0) -> np.ndarray:
        """min 1/2 u'Hu + (h (x0 - goal))'u  s.t. the enabled constraint blocks on x_ = T x0 + S u
        (absolute coordinates).  Returns u(0); side-effect attributes as in the reference (:340-347)."""
        x0 = np.asarray(x0, dtype=float)
        u, cost = self._solve_one(x0, self.goal)
        x_abs = self.T @ x0 + self.S @ u
        x0 = x0 - self.goal
        self.x_horizon = (self.T @ x0 + self.S @ u).reshape(-1, self.nx) + self.goal
        self.u_horizon = u.reshape(-1, 2)
        x_N = x_abs[-4:]                               # absolute x(N): the reference's quirk (:344-345)
        self.terminal_cost = x_N @ self.P @ x_N
        u0 = u[:self.nu]
        self.stage_cost = x0 @ self.Q @ x0 + u0 @ self.R @ u0
        self.cost = cost
        return u[:self.nu]


_C_XYV = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
_L_OBSERVER = np.array([[3.00000000e-01, 2.00284502e-16, 2.00000000e-01],
                        [3.48682242e+00, 1.90000000e+00, 3.86733776e-01],
                        [1.74341121e+00, 8.00000000e-01, 1.93366888e-01],
                        [-3.36822969e-16, -2.07305381e-16, 3.00000000e-01]])


def solve_target_selection(A, B, C, Q, R, d_state, y_target, Ax_ineq, bx_ineq, u_upper, u_lower):
    """Optimal target selection (reference :413-437): the 6-variable QP

        min x'Qx + u'Ru  s.t.  (I - A) x - B u = d_state,  C x = y_target,  state / input bounds.

    The equality block has 7 rows for 6 unknowns; for every equilibrium goal it is consistent and
    has full column rank, so the feasible set is the single point returned here (least-squares
    solve, residual and inequality checks asserted).  Otherwise a null-space QP is solved by
    scipy SLSQP on the host (6 variables; not a hot path - the reference re-solves it each step
    although it is constant).
    """
    nx, nu = B.shape
    E = np.block([[np.eye(nx) - A, -B], [C, np.zeros((C.shape[0], nu))]])
    rhs = np.hstack((d_state, y_target))
    sol, *_ = np.linalg.lstsq(E, rhs, rcond=None)
    rank = np.linalg.matrix_rank(E)
    consistent = np.linalg.norm(E @ sol - rhs) <= 1e-9 * max(1.0, np.linalg.norm(rhs))
    if not consistent:
        raise OutsideTheRegionOfAttractionError("target selection: no steady state reproduces y_goal")
    x_ref, u_ref = sol[:nx], sol[nx:]
    inside = np.all(Ax_ineq @ x_ref <= bx_ineq + 1e-9) and np.all(u_ref <= u_upper + 1e-9) \
        and np.all(u_ref >= u_lower - 1e-9)
    if rank == nx + nu:
        if not inside:
            raise OutsideTheRegionOfAttractionError("target selection: steady state violates bounds")
        return x_ref, u_ref
    from scipy.optimize import minimize
    W = np.block([[Q, np.zeros((nx, nu))], [np.zeros((nu, nx)), R]]).astype(float)
    cons = [{'type': 'eq', 'fun': lambda z: E @ z - rhs, 'jac': lambda z: E},
            {'type': 'ineq', 'fun': lambda z: bx_ineq - Ax_ineq @ z[:nx]},
            {'type': 'ineq', 'fun': lambda z: u_upper - z[nx:]},
            {'type': 'ineq', 'fun': lambda z: z[nx:] - u_lower}]
    res = minimize(lambda z: z @ W @ z, sol, jac=lambda z: 2 * W @ z, constraints=cons, method='SLSQP',
                   options={'ftol': 1e-14, 'maxiter': 200})
    if not res.success:
        raise OutsideTheRegionOfAttractionError("target selection failed: " + res.message)
    return res.x[:nx], res.x[nx:]


class MPCOutputFB(MPC):
    """Output feedback: y = C x with C selecting (x, y, v); Luenberger observer with the
    reference's hard-coded gain L (:401-404); the QP runs on the estimate."""

    def __init__(self, dt, N, lin_state, lin_input, init_state, terminal_constraint: bool = True,
                 input_constraint: bool = True, state_constraint: bool = True,
                 env: BaseEnv = None) -> None:
        super().__init__(dt, N, lin_state, lin_input, terminal_constraint, input_constraint,
                         state_constraint, env)
        self.C = _C_XYV.copy()
        self.y_goal = self.C @ env.goal
        self.L = _L_OBSERVER.copy()
        self.x_estimate = np.array(init_state)
        self.previous_u = np.zeros(2)
        self.check_constraints()

    def optimal_target_selection(self) -> tuple[np.ndarray, np.ndarray]:
        return solve_target_selection(self.A, self.B, self.C, self.Q, self.R, np.zeros(self.nx),
                                      self.y_goal, np.vstack(self.state_const_A),
                                      np.array(self.state_const_b), self.input_upper, self.input_lower)

    def luenberger_observer(self, y0):
        """x_hat+ = A x_hat + B u_prev + L (y - C x_hat)."""
        return self.A @ self.x_estimate + self.B @ self.previous_u + self.L @ (y0 - self.C @ self.x_estimate)

    def step(self, y0) -> np.ndarray:
        x0 = self.luenberger_observer(np.asarray(y0, dtype=float))
        self.x_estimate = x0
        x_ref, u_ref = self.optimal_target_selection()
        # The reference optimises over (variable - u_ref) and applies cost, constraints and the
        # return value to that expression (:461): the shift cancels, so u below is that expression.
        u, cost = self._solve_one(x0, x_ref)
        x0 = x0 - x_ref
        self.x_horizon = (self.T @ x0 + self.S @ u).reshape(-1, self.nx) + self.goal
        self.u_horizon = u.reshape(-1, 2)
        self.cost = cost
        self.previous_u = u[:self.nu]
        return u[:self.nu]


class MPCOutputFBWithDisturbance(MPC):
    """Output feedback with a scalar constant disturbance d: x+ = A x + B u + Bd d, y = C x + Cd d.

    Experimental in the reference (debug prints, hard-coded target [30, 1.5, 0, 0], :610-624); the
    same behaviour is kept, minus the prints.  The disturbance enters the prediction as a constant
    offset of the constraint right-hand sides, which the batched solver takes per sample.
    """

    def __init__(self, dt, N, lin_state, lin_input, init_state, terminal_constraint: bool = True,
                 input_constraint: bool = True, state_constraint: bool = True,
                 env: BaseEnv = None) -> None:
        super().__init__(dt, N, lin_state, lin_input, terminal_constraint, input_constraint,
                         state_constraint, env)
        self.C = _C_XYV.copy()
        self.Bd = np.array([1, 1, 1, 1])
        self.Cd = np.array([1, 1, 1])
        self.C_hat = np.hstack((self.C, self.Cd.reshape(-1, 1)))
        self.L1 = _L_OBSERVER.copy()
        self.L2 = np.array([1, 1, 1]) * 1
        self.A_hat = np.vstack((np.hstack((self.A, self.Bd.reshape(-1, 1))), np.array([0, 0, 0, 0, 1]).reshape(1, -1)))
        self.B_hat = np.vstack((self.B, [0, 0]))
        self.x_estimate = np.array(init_state)
        self.d_estimate = 0
        self.previous_u = np.zeros(2)
        self.y_goal = self.C @ env.goal
        self.check_constraints()

    def optimal_target_selection(self) -> tuple[np.ndarray, np.ndarray]:
        return solve_target_selection(self.A, self.B, self.C, self.Q, self.R, self.Bd * self.d_estimate,
                                      self.y_goal - self.Cd * self.d_estimate,
                                      np.vstack(self.state_const_A), np.array(self.state_const_b),
                                      self.input_upper, self.input_lower)

    def luenberger_observer(self, y0):
        innovation = y0 - self.C @ self.x_estimate - self.Cd * self.d_estimate
        x_new = self.A @ self.x_estimate + self.Bd * self.d_estimate + self.B @ self.previous_u + self.L1 @ innovation
        d_new = self.d_estimate + self.L2 @ innovation
        return x_new, d_new

    def disturbance_response(self) -> np.ndarray:
        """ABd with x_ = T x0 + S u + ABd d: block k is sum_{i<k} A^i Bd (reference :631-635)."""
        out = np.zeros((self.N + 1, self.nx))
        acc = np.zeros(self.nx)
        Ak_Bd = self.Bd.astype(float)
        for k in range(1, self.N + 1):
            acc = acc + Ak_Bd
            out[k] = acc
            Ak_Bd = self.A @ Ak_Bd
        return out.ravel()

    def step(self, y0) -> np.ndarray:
        x0, d0 = self.luenberger_observer(np.asarray(y0, dtype=float))
        x_ref, u_ref = np.array([30, 1.5, 0, 0]), np.zeros(2)
        self.x_estimate = x0
        self.d_estimate = d0
        res = self.batch_solver().solve_host(x0[None, :], x_ref=x_ref, want_u_full=True,
                                             c=np.array([float(self.d_estimate)]))
        if res.status[0] == 1:
            raise OutsideTheRegionOfAttractionError
        if res.status[0] != 0:
            raise RuntimeError(f"QP solver stopped with status {int(res.status[0])}")
        u = res.u_full[0]
        x0 = x0 - x_ref
        self.x_horizon = (self.T @ x0 + self.S @ u).reshape(-1, self.nx) + self.goal
        self.u_horizon = u.reshape(-1, 2)
        self.cost = float(res.objective[0])
        self.previous_u = u[:self.nu]
        return u[:self.nu]
