import os, sys
sys.path.insert(0, '/root/repo')
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
for N in (10,):
    c = _controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N)
    for ce in (12, 14, 16, 20):
        bq = BatchQP.from_controller(c, check_every=ce)
        bq.solve(x0)
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = bq.solve(x0); e1.record(); e1.synchronize(); ms.append(e0.elapsed_time(e1))
        it, la = bq.last_stats()
        print(f"N={N} check_every {ce}: {min(ms):.2f} ms, mean iters {it/x0.shape[1]:.2f}, undecided {(out['status']==2).sum().item()}", flush=True)
