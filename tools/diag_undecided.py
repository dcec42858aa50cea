"""Diagnostic: the samples of a config grid that end undecided (status 2) - where are they, what is their exact LP slack?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
from oracle import carmpc_oracle as orc
cases = [("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20, "RoadOneCarEnv_29.9_1.5_0_0.npy"), ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 10, "RoadOneCarEnv_29.9_1.5_0_0.npy"),
         ("RoadMultipleCarsEnv", None, 20, "RoadMultipleCarsEnv_30_1.5_0_0.npy")]
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
for env, goal, N, fx in cases:
    c = _controller(env, goal, N)
    bq = BatchQP.from_controller(c)
    xs = x0.clone()
    if goal is None:
        xs[0] += 0.1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = bq.solve(xs)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    und = np.flatnonzero(st == 2)
    print(f"== {env} N={N}: {dt*1e3:.1f} ms, undecided {len(und)}, stats {bq.polish_stats()}")
    oq = orc.CondensedQP(env, N, np.load(os.path.join(ROOT, "terminal_sets", fx)))
    pts = xs[:, torch.from_numpy(und[:60]).cuda()].cpu().numpy().T
    if len(pts):
        feas, slack = orc.qp_feasible_lp(oq, pts)
        for k in range(len(pts)):
            print(f"   {und[k]} x0={np.array2string(pts[k], precision=5)} iters={it[und[k]]} lp_slack={slack[k]:.3e}")
