"""Where an ADMM iteration's time goes: fixed 50 iterations per sample, parts switched off (timing only)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
c = _controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20)
bq = BatchQP.from_controller(c, eps_abs=0.0, eps_rel=0.0, eps_prim_inf=1e30, max_iter=50, polish=0)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = bq.solve(x0); torch.cuda.synchronize(); dt = time.perf_counter() - t0
it, la = bq.last_stats()
print(f"flags={os.environ.get('CARMPC_ADMM_DEBUG', '0')} wide={os.environ.get('CARMPC_ADMM_WIDE', '0')}: {dt*1e3:.2f} ms for {it/1e6:.1f} iterations/sample, launches {la}")
