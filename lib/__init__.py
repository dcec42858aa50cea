"""Import alias: the reference's module paths (``from lib.mpc import MPCStateFB``, examples/run_MPCStateFB.py:4-9 of
ahmad12hamdan99/CarMPC) resolve to the B200 implementations in ``carmpc_b200.lib``.

``lib.X`` *is* ``carmpc_b200.lib.X`` (the same module object, registered under both names), so class identities,
module-level settings such as ``lib.terminal_set.TERMINAL_SET_DIR`` and ``isinstance`` checks agree whichever path a
caller imports.  Plotting (``lib.visualize_state``) is out of scope (last section): importing it raises an
``ImportError`` that says so instead of silently drawing nothing.
"""
import importlib
import sys

_MODULES = ("configuration", "simulator", "environments", "matrix_gen", "in_adm_set", "polytope_ops", "mpc",
            "terminal_set")

for _name in _MODULES:
    _mod = importlib.import_module("carmpc_b200.lib." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod

del importlib, sys, _name, _mod
