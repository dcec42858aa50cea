import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
bq = BatchQP.from_controller(_controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N))
out = bq.solve(x0)
it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy()
for name, m in (("solved", st == 0), ("infeasible", st == 1), ("maxiter", st == 2)):
    if m.sum():
        v = it[m]
        print(name, m.sum(), "mean %.1f" % v.mean(), "pct 50/90/99/99.9/99.99/max", [int(np.percentile(v, p)) for p in (50, 90, 99, 99.9, 99.99)], v.max())
for thr in (100, 150, 200, 300, 500, 1000, 2000):
    print("iters >", thr, ":", int((it > thr).sum()), " sum of iterations above:", int(np.maximum(it - thr, 0).sum()))
print("total iterations", it.sum())
