"""Road environments: the extra state-constraint rows ``A x <= b``, the goal, a name and plot limits.

Drop-in for the reference's ``lib/environments.py`` (``BaseEnv`` :10-48, ``RoadEnv`` :52-78, ``RoadOneCarEnv``
:81-115, ``RoadMultipleCarsEnv`` :118-171): same classes and attributes (``constraints_A``, ``constraints_b``,
``goal``, ``name``, ``lim``, ``handles``, ``set_goal``, ``set_lim``, ``check_constraints``, ``plot``).  The numbers are
what the controllers and the terminal-set construction consume; they are declared as data below.  matplotlib is a
plotting-only dependency and is imported inside the drawing code, so the batch-evaluation path never needs it.
"""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np

from .simulator import CarTrailerDimension

GREY = (0.6, 0.6, 0.6)
YELLOW = (0.91, 0.8, 0.18)
PURPLE = (0.78, 0.18, 0.91)


def _plt():
    import matplotlib.pyplot as plt          # lazy: drawing only
    return plt


class BaseEnv:
    #: rows (a, b) of ``a . [x, y, psi, v] <= b`` this class adds on top of its parents'
    ROWS: Sequence[tuple] = ()
    #: legend entries (colour, label) this class adds
    LEGEND: Sequence[tuple] = ()
    NAME = 'BaseEnv'
    GOAL = None
    LIM = None

    def __init__(self):
        self.constraints_A, self.constraints_b = [], []
        for klass in reversed(type(self).__mro__):                    # parents first, as the reference appends
            for a, b in klass.__dict__.get('ROWS', ()):
                self.constraints_A.append(list(a))
                self.constraints_b.append(b)
        self.name = type(self).NAME
        self.goal = None if type(self).GOAL is None else list(type(self).GOAL)
        self.lim = None if type(self).LIM is None else list(type(self).LIM)
        self.check_constraints()

    # ---- data ---------------------------------------------------------------------------------
    def set_goal(self, goal: Union[list, np.ndarray, tuple]):
        self.goal = goal

    def set_lim(self, lim):
        """Axis ranges of the plot: ``[(x_min, x_max), (y_min, y_max)]``."""
        self.lim = lim

    def check_constraints(self):
        assert len(self.constraints_A) == len(self.constraints_b), \
            f"{len(self.constraints_A)} constraint rows but {len(self.constraints_b)} right-hand sides: " \
            "the two lists must have the same length."

    # ---- drawing -------------------------------------------------------------------------------
    @property
    def handles(self):
        """Legend handles: the goal star, then one patch per LEGEND entry from the base class down."""
        import matplotlib.lines as mlines
        import matplotlib.patches as patches
        out = [mlines.Line2D([], [], color='lime', marker='*', linestyle='None', markersize=10, label='goal')]
        for klass in reversed(type(self).__mro__):
            out += [patches.Patch(color=c, label=l) for c, l in klass.__dict__.get('LEGEND', ())]
        return out

    def plot(self) -> None:
        self.x_lim, self.y_lim = self.lim
        _plt().scatter(self.goal[0], self.goal[1], marker="*", c='lime')

    def _parked_cars(self, first_corner, count, colour):
        import matplotlib.patches as patches
        dim = CarTrailerDimension
        ax = _plt().gca()
        for k in range(count):
            corner = (first_corner[0] + k * 1.25 * dim.car_length, first_corner[1])
            ax.add_patch(patches.Rectangle(corner, dim.car_length, dim.car_width, edgecolor='none', facecolor=colour))


class RoadEnv(BaseEnv):
    """A straight road: ``-3 <= y <= 3``."""
    ROWS = (((0, 1, 0, 0), 3), ((0, -1, 0, 0), 3))
    LEGEND = ((GREY, 'Road boundaries'),)
    NAME = 'RoadEnv'
    GOAL = (30, 1.5, 0, 0)
    LIM = ((-10, 50), (-10, 10))
    y_lower, y_upper = -3, 3

    def plot(self) -> None:
        super().plot()
        plt = _plt()
        plt.fill_between(list(self.x_lim), self.y_upper, self.y_lim[1], color=GREY)
        plt.fill_between(list(self.x_lim), self.y_lower, self.y_lim[0], color=GREY)
        plt.plot(list(self.x_lim), [0, 0], color=(0, 0, 0), linestyle=(0, (5, 10)))


class RoadOneCarEnv(RoadEnv):
    """The road with a queue of cars ahead: additionally ``x <= 30``."""
    ROWS = (((1, 0, 0, 0), 30),)
    LEGEND = ((YELLOW, 'Obstacles'),)
    NAME = 'RoadOneCarEnv'
    GOAL = (29.9, -1.5, 0, 0)
    X_STOP = 30

    def plot(self) -> None:
        super().plot()
        plt = _plt()
        plt.vlines(self.X_STOP, self.y_lim[0], self.y_lim[1], color=YELLOW, linestyle='--')
        plt.fill_between([self.X_STOP, self.x_lim[1]], self.y_lim[0], self.y_lim[1], color=YELLOW + (0.1,))
        dim = CarTrailerDimension
        self._parked_cars((self.X_STOP + dim.car_length - dim.l12, -1.5 - dim.car_width / 2), 3, YELLOW)


class RoadMultipleCarsEnv(RoadEnv):
    """The road with two queues: ``y <= 0.25 x - 2`` and ``y >= 0.25 x - 6.25`` (rows that couple x and y)."""
    ROWS = (((-0.25, 1, 0, 0), -2), ((0.25, -1, 0, 0), 6.25))
    LEGEND = ((YELLOW, 'Obstacles'), (PURPLE, 'Obstacles'))
    NAME = 'RoadMultipleCarsEnv'
    GOAL = (30, 1.5, 0, 0)

    def _boundary(self, row, colour, fill_to):
        plt = _plt()
        a, b = self.constraints_A[row], self.constraints_b[row]
        xs = np.linspace(*self.x_lim, 2)
        ys = (b - a[0] * xs) / a[1]
        plt.plot(xs, ys, color=colour, linestyle='--')
        plt.fill_between(xs, ys, fill_to, color=colour + (0.1,))

    def plot(self) -> None:
        super().plot()
        dim = CarTrailerDimension
        self._boundary(-2, YELLOW, self.y_lim[1])
        self._parked_cars((-1 - 2 * 1.25 * dim.car_length, 1.5 - dim.car_width / 2), 3, YELLOW)
        self._boundary(-1, PURPLE, self.y_lim[0])
        self._parked_cars((30, -1.5 - dim.car_width / 2), 4, PURPLE)
