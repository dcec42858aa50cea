"""One seeded config-3 map solve inside a cudaProfilerStart/Stop range (ncu --profile-from-start off):
python tools/prof_seeded.py [block] [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid, lattice_seeds

block = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3x8x1x1").split("x"))
axes = config3_axes()
x0 = torch.stack(materialise_grid(axes, device="cuda")).contiguous()
seed = torch.from_numpy(lattice_seeds([len(a) for a in axes], block=block)).cuda()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bq = BatchQP.from_controller(_controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N))
for _ in range(2):
    bq.solve(x0, seed=seed)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = bq.solve(x0, seed=seed)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("seeded", out["seeded"], bq.polish_stats())
