"""Whole cold solve against the ADMM stopping tolerance (the float64 polish sets the accuracy either way): python tools/sweep_eps.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
for N in (80, 40):
    c = _controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N)
    for eps in (3e-3, 2e-3, 1.5e-3, 1e-3):
        bq = BatchQP.from_controller(c, eps_abs=eps, eps_rel=eps)
        bq.solve(x0)
        ms = []
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = bq.solve(x0); e1.record(); e1.synchronize(); ms.append(e0.elapsed_time(e1))
        it, la = bq.last_stats(); ps = bq.polish_stats()
        print(f"N={N} eps {eps:g}: {min(ms):.2f} ms, mean iters {it/x0.shape[1]:.2f}, certified after rounds {ps['certified_after_rounds'][:4]}, handed back {ps['handed_to_admm']}", flush=True)
