"""Tensor-core ADMM kernel against the FFMA kernel on the same config-3 states: python tools/tc_try.py [N] [states] [reps]."""
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
grid = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
if B < grid.shape[1]:
    idx = torch.randperm(grid.shape[1], generator=torch.Generator().manual_seed(1))[:B].cuda()
    x0 = grid[:, idx].contiguous()
else:
    x0 = grid
bq = BatchQP.from_controller(_controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N))
print("tensor form:", bq.tensor_mode())
res = {}
for mode in (0, 2):
    bq.tensor_mode(mode)
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = bq.solve(x0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        it, la = bq.last_stats()
        st = out["status"]
        info = bq.tensor_mode()
        print(f"mode {mode} N={N} B={x0.shape[1]} rep {r}: {dt * 1e3:.2f} ms, {x0.shape[1] / dt:.3e} QP/s, mean iters {it / x0.shape[1]:.2f}, "
              f"launches {la}, feasible {(st == 0).float().mean().item():.4f}, undecided {(st == 2).sum().item()}, "
              f"tc samples {info['samples_last_solve']}", flush=True)
    res[mode] = {k: v.clone() for k, v in out.items() if torch.is_tensor(v)}
    print("   polish:", bq.polish_stats()["certified_after_rounds"], "handed back", bq.polish_stats()["handed_to_admm"], flush=True)
bq.tensor_mode(3)
bq.solve(x0)
info = bq.tensor_mode()
cyc = info["cycles"]
names = ["mma: round total", "mma: wait A", "mma: wait B", "cmp: wait x~", "cmp: wait z^", "cmp: wait A stage", "cmp: retire/refill",
         "cmp: round total", "rounds", "tma: wait stage", "cmp: proxy fence", "cmp: fence + syncwarp"]
rounds = max(cyc[8], 1)
print("cycle counters per round (one round = check_every iterations of a 128-sample tile):")
for nm, v in zip(names, cyc):
    print(f"   {nm:22s} {(v if nm == 'rounds' else v / rounds):12.0f}")
bq.tensor_mode(1)
a, b = res[0], res[2]
same = (a["status"] == b["status"])
print("status equal:", same.float().mean().item(), "differences:", (~same).sum().item())
ok = (a["status"] == 0) & (b["status"] == 0)
du = (a["u0"] - b["u0"]).abs()[:, ok]
print("max |u0 diff| on commonly solved:", du.max().item() if ok.any() else None)
if "objective" in a:
    do = ((a["objective"] - b["objective"]).abs() / a["objective"].abs().clamp_min(1.0))[ok]
    print("max rel objective diff:", do.max().item())
