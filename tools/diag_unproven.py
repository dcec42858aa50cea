"""Diagnostic: how many samples leave the final polish without a KKT certificate (run on the GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
for env, goal, N in (("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20), ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 10), ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 40),
                     ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 80), ("RoadMultipleCarsEnv", None, 20), ("RoadEnv", None, 20)):
    bq = BatchQP.from_controller(_controller(env, goal, N))
    xs = x0.clone()
    if goal is None:
        xs[0] += 0.1
    bq.solve(xs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = bq.solve(xs)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ps = bq.polish_stats()
    st = out["status"]
    print(f"{env} N={N}: {dt*1e3:.2f} ms, solved {(st==0).sum().item()} infeasible {(st==1).sum().item()} maxiter {(st==2).sum().item()} "
          f"handed_to_admm {ps['handed_to_admm']} final_uncertified {ps['final_polish_uncertified']} "
          f"settled_own_cert {ps['max_iter_settled_by_own_certificate']} infeasible_before_second {ps['infeasible_before_second_pass']}", flush=True)
