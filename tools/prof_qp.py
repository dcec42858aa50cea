"""One config-3 solve (10^6 horizon-N QPs) for profiling: python tools/prof_qp.py [N] [states] [repeats]."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()[:, :B].contiguous()
bq = BatchQP.from_controller(_controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N))
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = bq.solve(x0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it, la = bq.last_stats()
    st = out["status"]
    print(f"N={N} B={B} rep {r}: {dt * 1e3:.2f} ms, {B / dt:.3e} QP/s, mean iters {it / B:.1f}, launches {la}, "
          f"feasible {(st == 0).float().mean().item():.4f}, max_iter {(st == 2).sum().item()}")
if len(sys.argv) > 4:
    bad = torch.nonzero(st == 2).flatten()[:50].cpu().numpy()
    print("max_iter samples:", x0[:, bad].cpu().numpy().T.tolist())
