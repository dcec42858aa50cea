"""The example scripts (examples/: headless counterparts of the reference's examples/*.py plus the batched map) run end to
end on the GPU and give the known results."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")

pytestmark = pytest.mark.gpu


@pytest.fixture()
def example():
    sys.path.insert(0, EXAMPLES)
    from carmpc_b200.lib import terminal_set as ts
    saved = ts.TERMINAL_SET_DIR
    yield lambda name: importlib.import_module(name)
    ts.TERMINAL_SET_DIR = saved
    sys.path.remove(EXAMPLES)


def test_find_terminal_set_reproduces_the_shipped_fixture_and_its_grid(example, tmp_path, capsys):
    A, b, inside = example("find_terminal_set").main(["--save", str(tmp_path)])
    out = capsys.readouterr().out
    assert "reproduced" in out
    assert inside.sum() == 434 and inside.reshape(6, -1).sum(axis=1).tolist() == [49, 70, 84, 91, 91, 49]
    saved = np.load(os.path.join(str(tmp_path), "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    assert saved.shape == (len(b), 5) and np.array_equal(saved[:, :4], A)


def test_state_feedback_example_reaches_the_goal(example):
    states, inputs, costs, reached = example("run_MPCStateFB").main([])
    assert reached and len(costs) > 10
    assert np.abs(inputs[:, 0]).max() <= 2.0 + 1e-6 and np.abs(inputs[:, 1]).max() <= np.pi / 8 + 1e-6
    dist = np.abs(states[:, 0] - 30.0)
    assert np.all(np.diff(dist) <= 1e-9), "the car must approach the goal position monotonically along the road"


def test_output_feedback_and_lqr_examples_reach_the_goal(example):
    state, reached, err = example("run_MPCOutputFB").main([])
    assert reached and err[-1] < 0.05
    state, reached = example("run_LQR").main([])
    assert abs(state[0] - 30.0) <= 0.2 and abs(state[3]) <= 0.1     # the unconstrained law parks the car along the road


def test_roa_map_example(example):
    A, b, flags, out = example("roa_map").main(["--points", "24", "24", "3", "4", "--runs", "2000", "--steps", "60"])
    assert 0.3 < flags.mean() < 0.99 and len(b) >= 8
    fail = out["fail_step"].cpu().numpy()
    assert (fail < 0).sum() > 200
