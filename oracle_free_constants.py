"""Observer constants of the output-feedback controller (lib/mpc.py:387-404) for the benchmark, taken from the
product's own controller class so that bench.py's GPU arm does not touch oracle/."""
from carmpc_b200.lib.mpc import _C_XYV as C_OUT, _L_OBSERVER as L_OBS   # noqa: F401
