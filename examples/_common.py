"""Shared by the example scripts: repo root on sys.path and the shipped terminal sets as the controller's set directory
(the reference resolves '../terminal_sets/' from CWD = examples/, lib/mpc.py:98; here the scripts run from anywhere)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from carmpc_b200.lib import terminal_set as _ts          # noqa: E402

_ts.TERMINAL_SET_DIR = os.path.join(ROOT, "terminal_sets") + os.sep


def make_env(name: str, goal=None):
    from carmpc_b200.lib import environments
    env = getattr(environments, name)()
    if goal is not None:
        env.set_goal(goal)
    return env
