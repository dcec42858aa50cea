"""Scan time vs sample count and staging geometry on one GPU (back-to-back launches, CUDA events)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from carmpc_b200.batch import TerminalSetEvaluator, RolloutEvaluator
from carmpc_b200.grids import config2_axes, materialise_grid
from carmpc_b200.lib.environments import RoadMultipleCarsEnv
Ab = np.load(os.path.join(ROOT, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
x, y, psi, v = materialise_grid(config2_axes(), device="cuda")
geos = [(256, 3, 1), (128, 3, 1), (128, 4, 1), (128, 2, 2), (128, 3, 2), (256, 2, 1), (128, 2, 1), (256, 2, 2)]
which = sys.argv[1] if len(sys.argv) > 1 else "hrep"
def timed(ev, n, off=0, steps=20):
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda"); cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    a = [t[off:off + n] for t in (x, y, psi, v)]
    for _ in range(3): ev.contains_bits(*a, bits=bits, count=cnt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): ev.contains_bits(*a, bits=bits, count=cnt)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / steps, int(cnt.item())
for g in geos:
    ev = TerminalSetEvaluator(Ab) if which == "hrep" else RolloutEvaluator.from_env(RoadMultipleCarsEnv(), 16)
    ev.set_staging(*g)
    row = []
    for n in (12_500_992, 25_000_960, 50_000_896, 100_000_000):
        ms, c = timed(ev, n)
        row.append(f"{n/1e6:.1f}M: {ms*1e3:.1f} us ({n*32.125/ms/1e6:.0f} GB/s)")
    ms8, _ = timed(ev, 12_500_992, off=87_500_000 - 87_500_000 % 1024)
    print(which, g, " | ".join(row), f"| last 1/8: {ms8*1e3:.1f} us", flush=True)
