"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol check, gloo sharding (CPU only).
``-m gpu``: parity of the CUDA path against the oracle, through the C ABI (needs a B200).
"""
import os
import sys

import numpy as np
import pytest

# tests/test_sharded_gpu.py runs the ranks of a group as streams of ONE process; their exchange kernels wait for each
# other, so no two of those streams may share a hardware queue (the default is 8 queues).  Read at context creation.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
TERMINAL_SETS = os.path.join(ROOT, "terminal_sets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) when selected with -m gpu on a machine without a device; without an
    explicit -m selection on a CPU-only machine they are skipped."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


FIXTURES = {   # file -> (environment class name, goal)
    "RoadOneCarEnv_29.9_1.5_0_0.npy": ("RoadOneCarEnv", [29.9, 1.5, 0, 0]),
    "RoadOneCarEnv_29.9_-1.5_0_0.npy": ("RoadOneCarEnv", [29.9, -1.5, 0, 0]),
    "RoadMultipleCarsEnv_30_1.5_0_0.npy": ("RoadMultipleCarsEnv", [30, 1.5, 0, 0]),
    "RoadEnv_30_1.5_0_0.npy": ("RoadEnv", [30, 1.5, 0, 0]),
    "RoadEnv_30_0_0_0.npy": ("RoadEnv", [30, 0, 0, 0]),
}
K_STAR = {"RoadOneCarEnv": 20, "RoadMultipleCarsEnv": 16, "RoadEnv": 11}     # SURVEY 3.1 [probe]


def make_env(name, goal=None):
    from carmpc_b200.lib import environments
    env = getattr(environments, name)()
    if goal is not None:
        env.set_goal(goal)
    return env


def make_controller(env, N=20, cls="MPCStateFB", **kw):
    from carmpc_b200.lib import mpc
    from carmpc_b200.lib.configuration import DT_CONTROL, LINEARIZE_STATE, LINEARIZE_INPUT
    cwd = os.getcwd()
    os.chdir(os.path.join(ROOT, "tests"))            # '../terminal_sets/' is relative, as in the reference
    try:
        return getattr(mpc, cls)(dt=DT_CONTROL, N=N, lin_state=LINEARIZE_STATE, lin_input=LINEARIZE_INPUT, env=env, **kw)
    finally:
        os.chdir(cwd)
