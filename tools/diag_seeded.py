"""Diagnostic: where do the cold and the seeded solve disagree on a test grid?  python tools/diag_seeded.py ENV N"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import make_env, make_controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import lattice_seeds, materialise_grid
env_name, N = sys.argv[1], int(sys.argv[2])
c = make_controller(make_env(env_name, [29.9, 1.5, 0, 0] if env_name == "RoadOneCarEnv" else None), N)
bq = BatchQP.from_controller(c)
g = np.array(c.goal, dtype=float)
axes = [np.linspace(g[0] - 20.0, g[0] + 0.5, 24), np.linspace(-3.0, 3.0, 40), np.linspace(-0.3, 0.3, 3), np.linspace(-1.0, 4.0, 4)]
x0 = torch.stack(materialise_grid(axes, device="cuda")).contiguous()
cold = bq.solve(x0)
seed_np = lattice_seeds([len(a) for a in axes], block=(2, 8, 1, 1))
warm = bq.solve(x0, seed=torch.from_numpy(seed_np).cuda())
sa, sb = cold["status"].cpu().numpy(), warm["status"].cpu().numpy()
ia, ib = cold["iters"].cpu().numpy(), warm["iters"].cpu().numpy()
d = np.flatnonzero(sa != sb)
print("mismatches", d, "cold status/iters", sa[d], ia[d], "seeded status/iters", sb[d], ib[d], "seed", seed_np[d], "seed status", sb[seed_np[d]])
print("x0", x0.cpu().numpy().T[d])
print("status counts cold", np.bincount(sa, minlength=3), "seeded", np.bincount(sb, minlength=3))
one = bq.solve(x0[:, d].contiguous())
print("solved alone: status", one["status"].cpu().numpy(), "iters", one["iters"].cpu().numpy())
