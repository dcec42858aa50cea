import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_qp_gpu import _setup
from oracle import carmpc_oracle as orc
fx = np.load(os.path.join(ROOT, "tests", "golden", "closed_loop_config4.npz"))
c, bq, oq = _setup("RoadEnv", 20)
g = np.array(c.goal, float)
x_init = fx["x_init"]
x1 = orc.plant_step(x_init, np.zeros((len(x_init), 2)))
r = bq.solve_host(x1[177:178]); print("alone:", r.status, r.iters, r.u0, bq.polish_stats())
r = bq.solve_host(x1); print("batch of 200: status[177] =", r.status[177], "iters", r.iters[177], "n_infeasible", (r.status == 1).sum(), bq.polish_stats())
ue, obje, ste, pol, slack = orc.qp_solve_exact(oq, x1, g)
bad = np.flatnonzero(np.where(r.status == 0, 0, 1) != ste)
print("flag mismatches in the batch:", bad, slack[bad])
out = bq.closed_loop(torch.from_numpy(np.ascontiguousarray(x_init.T)).cuda(), 1, c.A, c.B)
print("closed loop 1 step: fail[177] =", out["fail_step"][177].item(), "n_fail", (out["fail_step"] >= 0).sum().item(), bq.polish_stats())
for R in (178, 180, 192, 200):
    out = bq.closed_loop(torch.from_numpy(np.ascontiguousarray(x_init[:R].T)).cuda(), 1, c.A, c.B)
    f = out["fail_step"].cpu().numpy()
    print(R, "runs: fail[177] =", f[177], "mismatch vs oracle:", np.flatnonzero((f >= 0) != (ste[:R] == 1)))
