"""The reference's import lines work unchanged ("same module paths")."""
import importlib

import numpy as np
import pytest


@pytest.mark.parametrize("name", ["mpc", "terminal_set", "matrix_gen", "in_adm_set", "simulator", "environments",
                                  "configuration"])
def test_lib_module_is_the_b200_module(name):
    ours = importlib.import_module("carmpc_b200.lib." + name)
    alias = importlib.import_module("lib." + name)
    assert alias is ours


def test_reference_import_lines():
    # examples/run_MPCStateFB.py:4-9, run_MPCOutputFB.py:4-9, find_terminal_set.py:1-4 of the reference
    from lib.mpc import MPCStateFB, MPCOutputFB, MPC, OutsideTheRegionOfAttractionError      # noqa: F401
    from lib.simulator import CarSimulator, CarTrailerDimension                               # noqa: F401
    from lib.environments import RoadEnv, RoadOneCarEnv, RoadMultipleCarsEnv                   # noqa: F401
    from lib.configuration import DT_CONTROL, DT_SIMULATION, N, LINEARIZE_STATE, LINEARIZE_INPUT
    from lib.terminal_set import calc_terminal_set, compute_terminal_set, visualise_set       # noqa: F401
    from lib.matrix_gen import predmod, costgen
    from lib.in_adm_set import algorithm_1, algorithm_2                                       # noqa: F401
    A, B = MPC.discretized_model(*MPC.linearized_model(np.array(LINEARIZE_STATE), np.array(LINEARIZE_INPUT)), DT_CONTROL)
    T, S = predmod(A, B, 3)
    assert T.shape == (16, 4) and S.shape == (16, 6)
    H, h, _ = costgen(np.eye(4), np.eye(2), np.eye(4), T, S, 4)
    assert H.shape == (6, 6) and h.shape == (6, 4)
    assert DT_SIMULATION > 0 and N == 20
