"""Scan time of one rank's cyclic shard (every 8th 1024-sample group of the config-2 grid, compacted) as a plain scan:
separates the cost of the samples themselves from the cost of the sharded sink.  python tools/shard_scan_time.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from carmpc_b200.batch import TerminalSetEvaluator
from carmpc_b200.grids import config2_axes, materialise_grid
Ab = np.load(os.path.join(ROOT, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
cols = materialise_grid(config2_axes(), device="cuda")
n = len(cols[0])
def timed(ev, a, steps=50):
    m = len(a[0])
    bits = torch.empty((m + 31) // 32, dtype=torch.int32, device="cuda"); cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(5): ev.contains_bits(*a, bits=bits, count=cnt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): ev.contains_bits(*a, bits=bits, count=cnt)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3
ev = TerminalSetEvaluator(Ab)
ev.contains_bits(*cols)                                   # tunes the row order on the whole grid
print(f"whole grid: {timed(ev, cols, 20):.1f} us")
for world in (2, 4, 8):
    g = n // 1024
    for r in (0, world - 1):
        sel = (torch.arange(g, device="cuda") % world == r)
        shard = [c[: g * 1024].view(g, 1024)[sel].reshape(-1).contiguous() for c in cols]
        print(f"world {world} rank {r}: {len(shard[0])} samples, plain scan of the compact shard {timed(ev, shard):.1f} us "
              f"(whole-grid rate would give {len(shard[0]) / n * 519:.1f} us)")
    first = [c[: n // world].contiguous() for c in cols]
    print(f"world {world}: contiguous first 1/{world}: {timed(ev, first):.1f} us")
