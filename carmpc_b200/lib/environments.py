"""Road environments: extra state constraint rows, goal, name and plot limits.

Mirror of the reference's ``lib/environments.py`` (``BaseEnv`` :10-48, ``RoadEnv`` :52-78,
``RoadOneCarEnv`` :81-115, ``RoadMultipleCarsEnv`` :118-171).  The numerical content (rows of
``A x <= b``, goals, names, limits) is identical; matplotlib is imported only inside ``plot`` /
``handles`` because it is a plotting dependency and not needed by the batch evaluation path.
"""
from typing import Union

import numpy as np

from .simulator import CarTrailerDimension

_ROAD_GREY = (0.6, 0.6, 0.6)
_YELLOW = (0.91, 0.8, 0.18)
_PURPLE = (0.78, 0.18, 0.91)


def _pyplot():
    import matplotlib.pyplot as plt     # lazy: plotting only
    return plt


class BaseEnv:
    #: (colour, label) of the legend patches each class contributes; materialised lazily
    _legend_patches: list = []

    def __init__(self):
        self.constraints_A = []
        self.constraints_b = []
        self.goal = None
        self.lim = None
        self.name = 'BaseEnv'
        self.check_constraints()

    @property
    def handles(self):
        """Legend handles (goal star + per-class patches), built on first use."""
        import matplotlib.lines as mlines
        import matplotlib.patches as patches
        star = mlines.Line2D([], [], color='lime', marker='*', linestyle='None', markersize=10,
                             label='goal')
        out = [star]
        for klass in reversed(type(self).__mro__):
            for colour, label in klass.__dict__.get('_legend_patches', []):
                out.append(patches.Patch(color=colour, label=label))
        return out

    def set_goal(self, goal: Union[list, np.ndarray, tuple]):
        self.goal = goal

    def set_lim(self, lim):
        """Axis ranges of the plot: [(x_min, x_max), (y_min, y_max)]."""
        self.lim = lim

    def check_constraints(self):
        assert len(self.constraints_A) == len(self.constraints_b), \
            f"The constraints array (len(A) = {len(self.constraints_A)} and len(B) = {len(self.constraints_b)})" + \
            "do not have the same length. This will produced unpredictable behaviour."

    def plot(self) -> None:
        self.x_lim, self.y_lim = self.lim
        _pyplot().scatter(self.goal[0], self.goal[1], marker="*", c='lime')


class RoadEnv(BaseEnv):
    """Straight road: -3 <= y <= 3."""
    _legend_patches = [(_ROAD_GREY, 'Road boundaries')]

    def __init__(self):
        super().__init__()
        self.constraints_A += [[0, 1, 0, 0], [0, -1, 0, 0]]
        self.constraints_b += [3, 3]
        self.y_lower, self.y_upper = -3, 3
        self.lim = [(-10, 50), (-10, 10)]
        self.name = 'RoadEnv'
        self.check_constraints()
        self.goal = [30, 1.5, 0, 0]

    def plot(self) -> None:
        super().plot()
        plt = _pyplot()
        plt.fill_between([*self.x_lim], self.y_upper, self.y_lim[1], color=_ROAD_GREY)
        plt.fill_between([*self.x_lim], self.y_lower, self.y_lim[0], color=_ROAD_GREY)
        plt.plot(list(self.x_lim), [0, 0], color=(0, 0, 0), linestyle=(0, (5, 10)))


class RoadOneCarEnv(RoadEnv):
    """Road with a queue of cars ahead: additionally x <= 30."""
    _legend_patches = [(_YELLOW, 'Obstacles')]

    def __init__(self):
        super().__init__()
        self.constraints_A += [[1, 0, 0, 0]]
        self.constraints_b += [30]
        self.name = 'RoadOneCarEnv'
        self.check_constraints()
        self.goal = [29.9, -1.5, 0, 0]

    def plot(self) -> None:
        super().plot()
        import matplotlib.patches as patches
        plt = _pyplot()
        ax = plt.gca()
        x_upper = 30
        plt.vlines(x_upper, self.y_lim[0], self.y_lim[1], color=_YELLOW, linestyle='--')
        plt.fill_between([x_upper, self.x_lim[1]], self.y_lim[0], self.y_lim[1], color=_YELLOW + (0.1,))
        dim = CarTrailerDimension
        for i in range(3):
            corner = (30 + dim.car_length - dim.l12 + i * 1.25 * dim.car_length, -1.5 - dim.car_width / 2)
            ax.add_patch(patches.Rectangle(corner, dim.car_length, dim.car_width, edgecolor='none',
                                           facecolor=_YELLOW))


class RoadMultipleCarsEnv(RoadEnv):
    """Road with two queues: y <= 0.25 x - 2 and y >= 0.25 x - 6.25 (coupled x-y rows)."""
    _legend_patches = [(_YELLOW, 'Obstacles'), (_PURPLE, 'Obstacles')]

    def __init__(self):
        super().__init__()
        self.constraints_A += [[-0.25, 1, 0, 0], [0.25, -1, 0, 0]]
        self.constraints_b += [-2, 6.25]
        self.name = 'RoadMultipleCarsEnv'
        self.check_constraints()
        self.goal = [30, 1.5, 0, 0]

    def _line(self, k, colour, fill_to):
        plt = _pyplot()
        a, b = self.constraints_A[k], self.constraints_b[k]
        x = np.linspace(*self.x_lim, 2)
        y = (-a[0] * x + b) / a[1]
        plt.plot(x, y, color=colour, linestyle='--')
        plt.fill_between(x, y, fill_to, color=colour + (0.1,))

    def plot(self) -> None:
        super().plot()
        import matplotlib.patches as patches
        ax = _pyplot().gca()
        dim = CarTrailerDimension
        self._line(-2, _YELLOW, self.y_lim[1])
        for i in range(3):
            corner = (-1 - i * 1.25 * dim.car_length, 1.5 - dim.car_width / 2)
            ax.add_patch(patches.Rectangle(corner, dim.car_length, dim.car_width, edgecolor='none',
                                           facecolor=_YELLOW))
        self._line(-1, _PURPLE, self.y_lim[0])
        for i in range(4):
            corner = (30 + i * 1.25 * dim.car_length, -1.5 - dim.car_width / 2)
            ax.add_patch(patches.Rectangle(corner, dim.car_length, dim.car_width, edgecolor='none',
                                           facecolor=_PURPLE))
