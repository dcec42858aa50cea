"""How much of the config-3 ADMM time is the tail (a few slow samples keeping tiles busy) and how much the bulk?
Compares the normal solve with fixed-iteration runs (eps = 0: every feasible sample runs exactly max_iter iterations)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
ctl = _controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20)

def timed(bq, x):
    bq.solve(x)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); out = bq.solve(x); s1.record(); s1.synchronize()
        best = min(best, s0.elapsed_time(s1))
    return best, out

bq = BatchQP.from_controller(ctl)
ms, out = timed(bq, x0)
it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy()
print(f"normal solve: {ms:.2f} ms, mean iters {it.mean():.2f}, pct 50/90/99/99.9/max {[int(np.percentile(it, p)) for p in (50, 90, 99, 99.9)]} {it.max()}")
bq0 = BatchQP.from_controller(ctl, polish=0)
ms0, _ = timed(bq0, x0)
print(f"polish off (ADMM + pass-through): {ms0:.2f} ms")
feas = torch.from_numpy(st == 0).cuda()
xf = x0[:, feas].contiguous()
print("feasible states:", xf.shape[1])
for iters in (20, 40, 80):
    b = BatchQP.from_controller(ctl, polish=0, eps_abs=1e-30, eps_rel=1e-30, max_iter=iters)
    ms_i, o = timed(b, xf)
    total = int(o["iters"].sum().item())
    per = ms_i * 1e-3 / (total / 128 / 148)
    print(f"fixed {iters} iterations on feasible states: {ms_i:.2f} ms, total sample-iterations {total:.3e}, {per * 1e6:.2f} us per 128-sample tile-iteration")
# sorted by difficulty: the tail disappears when slow samples start first
order = torch.from_numpy(np.argsort(-it, kind="stable")).cuda()
xs = x0[:, order].contiguous()
ms_s, _ = timed(bq0, xs)
print(f"polish off, samples sorted slowest-first: {ms_s:.2f} ms")
