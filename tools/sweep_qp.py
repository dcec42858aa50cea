"""Option sweep of the batched QP on config 3 (development aid): python tools/sweep_qp.py"""
import itertools
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
if B < x0.shape[1]:
    x0 = x0[:, torch.arange(0, x0.shape[1], x0.shape[1] // B, device="cuda")[:B]].contiguous()
c = _controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N)
ref = None
grid = []
for eps in (1e-3, 1.5e-3, 2e-3, 2.5e-3, 4e-3):
    grid.append(dict(eps_abs=eps, eps_rel=eps))
for opts in [dict()] + grid:
    bq = BatchQP.from_controller(c, **opts)
    bq.solve(x0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = bq.solve(x0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    it, la = bq.last_stats()
    st = out["status"]
    u0 = out["u0"].nan_to_num(0.0)
    if ref is None:
        ref = (st.clone(), u0.clone())
    flag_diff = int(((st == 0) != (ref[0] == 0)).sum().item())
    both = (st == 0) & (ref[0] == 0)
    du = float((u0 - ref[1]).abs().max(0).values[both].max().item())
    print(f"{str(opts):45s} {dt * 1e3:7.2f} ms  {x0.shape[1] / dt:.3e} QP/s  iters {it / x0.shape[1]:6.1f}  launches {la} "
          f"maxit {(st == 2).sum().item():4d}  flag_diff {flag_diff:4d}  max|du0| {du:.1e}", flush=True)
    bq.close()
