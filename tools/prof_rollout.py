"""Rollout-form scan of the config-2 grid, for profiling: python tools/prof_rollout.py [repeats]."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from carmpc_b200.batch import RolloutEvaluator
from carmpc_b200.grids import config2_axes, materialise_grid
from carmpc_b200.lib.environments import RoadMultipleCarsEnv
x, y, psi, v = materialise_grid(config2_axes(), device="cuda")
ev = RolloutEvaluator.from_env(RoadMultipleCarsEnv(), 16)
n = len(x)
bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda"); cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    ev.contains_bits(x, y, psi, v, bits=bits, count=cnt)
torch.cuda.synchronize()
print("members", int(cnt.item()))
