"""Terminal-set construction and sampled membership evaluation.

Same public functions as the reference's ``lib/terminal_set.py``:

* ``compute_terminal_set`` (:23-73) - Gilbert-Tan maximal output-admissible set of
  x+ = A_k x under ``A_con x <= b_con``; the reference solves its LPs with cvxpy, here they go to
  ``scipy.optimize.linprog`` (HiGHS).  Host-side setup, seconds at most.
* ``calc_terminal_set`` (:135-213) - glue: builds the horizon-1 controller, intersects with the
  t = 0 input rows, removes redundant rows and writes ``<dir>/<env.name>_<goal...>.npy`` as
  float64 ``[A | b]`` rows (``A x <= b`` in absolute coordinates, unit-norm rows).
* ``visualise_set`` (:76-132) - samples the (x, y) mesh x 6 velocities at psi = 0 and tests
  ``A p <= b`` for every point.  The reference does this with a pure-Python triple loop (:107-113);
  here the whole grid goes through the CUDA membership kernel in one launch
  (``grid_membership``) and plotting is optional.

Batch entry points that have no counterpart in the reference (it evaluates one point at a time)
live in ``carmpc_b200.batch``.
"""
from __future__ import annotations

import itertools
import os
from typing import Optional

import numpy as np
from scipy.optimize import linprog

from .configuration import *          # noqa: F401,F403  (reference star-imports it, :12)
from .environments import BaseEnv
from . import polytope_ops as pc

#: where ``calc_terminal_set`` writes and ``MPC`` reads; the reference hard-codes this relative path
#: (``lib/terminal_set.py:210``, ``lib/mpc.py:98``), which only works with CWD = examples/.
TERMINAL_SET_DIR = '../terminal_sets/'


def compute_terminal_set(A_k: np.ndarray, A_con: np.ndarray, b_con: np.ndarray,
                         env: BaseEnv = None) -> tuple[Optional[pc.Polytope], Optional[int]]:
    """Maximal set invariant under x+ = A_k x inside {A_con x <= b_con}, around ``env.goal``.

    Returns ``(polytope, k)`` where the polytope stacks ``A_con A_k^t`` for t = 0..k, or
    ``(None, None)`` after 100 unsuccessful outer iterations (reference :71-73).
    The reference dereferences ``env.goal`` unconditionally (:42); ``env=None`` is accepted here and
    means the origin.
    """
    nx = A_k.shape[1]
    goal = np.zeros(nx) if env is None else np.array(env.goal, dtype=float)

    shifted = pc.Polytope(A_con, b_con).translation(-goal)
    A_con, b_con = shifted.A, shifted.b
    s = len(A_con)
    free = [(None, None)] * nx

    blocks = [A_con.copy()]                         # blocks[t] = A_con A_k^t
    for k in itertools.count(1):
        while len(blocks) < k + 2:
            blocks.append(blocks[-1] @ A_k)
        A_set = np.vstack(blocks[:k + 1])
        b_set = np.tile(b_con, k + 1)
        nxt = blocks[k + 1]
        admissible = True
        for i in range(s):
            res = linprog(-nxt[i], A_ub=A_set, b_ub=b_set, bounds=free, method="highs")
            # unbounded / failed LP: cvxpy reports +inf, which fails the <= 1e-3 test
            worst = -res.fun - b_con[i] if res.status == 0 else np.inf
            if not worst <= 1e-3:
                admissible = False
                break
        if admissible:
            return pc.Polytope(A_set, b_set).translation(goal), k
        if k > 100:
            print("Unsuccesful search: too many iterations. The goal might be outside the constraints.")
            return None, None


def grid_points(env: Optional[BaseEnv], extent: float = 25, steps: int = 100,
                v_range=None, psi: float = 0.0):
    """The sample grid of ``visualise_set`` (reference :96-111) as four SoA float64 arrays.

    Order: v slowest, then the mesh row i (y index), then the mesh column j (x index) - the order
    in which the reference's triple loop visits the points.
    """
    center = [0, 0, 0, 0] if env is None else env.goal
    v_range = np.arange(6) if v_range is None else np.asarray(v_range)
    xs = np.linspace(-extent + center[0], extent + center[0], steps)
    ys = np.linspace(-extent + center[1], extent + center[1], steps)
    xx, yy = np.meshgrid(xs, ys)
    nv = len(v_range)
    x = np.tile(xx.ravel(), nv)
    y = np.tile(yy.ravel(), nv)
    v = np.repeat(v_range.astype(float), xx.size)
    return x, y, np.full_like(x, psi), v


def grid_membership(A_constraints: np.ndarray, b_constraints: np.ndarray, env: Optional[BaseEnv],
                    extent: float = 25, steps: int = 100) -> tuple[np.ndarray, np.ndarray]:
    """Membership of every ``visualise_set`` grid point, evaluated on the GPU.

    Returns ``(points (n, 4), member (n,) bool)`` in the reference's visiting order.
    """
    from ..batch import TerminalSetEvaluator
    x, y, psi, v = grid_points(env, extent, steps)
    member = TerminalSetEvaluator(A_constraints, b_constraints).contains_host(x, y, psi, v)
    return np.stack((x, y, psi, v), axis=1), member


def visualise_set(A_constraints: np.ndarray, b_constraints: np.ndarray, env: BaseEnv,
                  extent: float = 25) -> None:
    """Project a set {A x <= b} on the x,y-plane for v = 0..5 m/s, psi = 0 (reference :76-132)."""
    halfspaces = np.hstack((A_constraints, np.expand_dims(b_constraints, axis=1)))
    A, b = halfspaces[..., :4], halfspaces[..., 4]
    points, member = grid_membership(A, b, env, extent)

    try:
        import matplotlib.pyplot as plt
        import matplotlib.patches as patches
        from scipy.spatial import ConvexHull
    except ImportError:
        print(f"{int(member.sum())} of {len(member)} grid points are inside the set "
              f"(matplotlib is not installed; nothing plotted).")
        return

    plt.figure()
    ax = plt.gca()
    blue = np.array([17, 73, 112]) / 255
    v_range = np.arange(6)
    for idx, v in enumerate(v_range):
        inside = points[member & (points[:, 3] == v)][:, :2]
        if len(inside) >= 3:
            hull = ConvexHull(inside)
            ax.add_patch(patches.Polygon(inside[hull.vertices], closed=True, edgecolor='none',
                                         facecolor=blue + (idx + 1) * 0.5 / len(v_range)))
    env.plot()
    plt.title(f"The terminal set projected on the x,y-plane \n for v = {v_range} m/s and " + r'$\psi$' + " = 0 rad")
    plt.xlim(*env.lim[0])
    plt.ylim(*env.lim[1])
    handles, _ = ax.get_legend_handles_labels()
    term_set = patches.Patch(color=blue + 0.25, label='Terminal set')
    ax.legend(loc='upper center', handles=handles + env.handles + [term_set],
              bbox_to_anchor=(0.5, -0.15), fancybox=True, shadow=True, ncol=3)
    ax.set_aspect('equal', 'box')
    plt.show()


def terminal_set_filename(env: BaseEnv) -> str:
    """``<name>_<g0>_<g1>_<g2>_<g3>.npy`` with ``str()`` of each goal element (reference :208)."""
    return env.name + ''.join('_' + str(s) for s in env.goal) + '.npy'


def lqr_closed_loop(env: BaseEnv):
    """(A_k, K, A_con, b_con, A_input K, b_input) of the horizon-1 controller (reference :145-161, 198-199)."""
    from .mpc import MPCStateFB
    controller = MPCStateFB(dt=DT_CONTROL, N=1, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT,
                            terminal_constraint=False, input_constraint=False, state_constraint=True,
                            env=env)
    A_k = controller.A + controller.B @ controller.K
    A_con, b_con = controller.state_constraint()
    A_con = A_con[..., 4:]                       # rows act on x(1); N = 1
    A_input, b_input = controller.input_constraint()
    return controller, A_k, A_con, b_con, A_input @ controller.K, b_input


def calc_terminal_set(env: BaseEnv, verbose: bool = False, save: bool = True) -> tuple[np.ndarray, np.ndarray]:
    """Terminal set of ``env``: invariant under the LQR law, inside the state constraints for all
    time and inside the input constraints at t = 0.  Writes the ``.npy`` (the reference ignores
    ``save`` and always writes, :210; here ``save=False`` really skips the write)."""
    controller, A_k, A_con, b_con, A_in, b_in = lqr_closed_loop(env)
    if verbose:
        ctrb = np.hstack([np.linalg.matrix_power(controller.A, i) @ controller.B for i in range(4)])
        print("============ Controller ============\n")
        print(f"A:\n{controller.A}\n\nB:\n{controller.B}\n\nK:\n{np.round(controller.K, 2)}\n\nA_k:\n{np.round(A_k, 2)}\n")
        print("============ Constraints ============\n")
        print(f"A {A_con.shape}:\n{A_con}\n\nb {b_con.shape}:\n{b_con}")
        print(f"\nThe rank of the controllability matrix is {np.linalg.matrix_rank(ctrb)} with n = {controller.nx}")
        print(f"eig(A)   = {np.linalg.eigvals(controller.A)}\neig(A_k) = {np.linalg.eigvals(A_k)}")

    p, _ = compute_terminal_set(A_k, A_con, b_con, env=env)
    if p is None:
        raise RuntimeError("no terminal set found within 100 iterations; is the goal inside the constraints?")
    p_input = pc.Polytope(A_in, b_in).translation(np.array(env.goal, dtype=float))
    p_terminal = pc.reduce(p.intersect(p_input))

    A_final, b_final = p_terminal.A, p_terminal.b
    halfspaces = np.hstack((A_final, np.expand_dims(b_final, axis=1)))
    filename = terminal_set_filename(env)
    if save:
        print(f"Filename to save: {filename}")
        os.makedirs(TERMINAL_SET_DIR, exist_ok=True)
        np.save(os.path.join(TERMINAL_SET_DIR, filename), halfspaces)
    print(f"Terminal set has {len(halfspaces)} constraints.")
    return A_final, b_final
