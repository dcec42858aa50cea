#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
python - <<'PY' 2>&1 | grep -v Using
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
for N in (10, 20):
    c = _controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N)
    base = None
    for rho in (0.0, 1.0, 2.0, 4.0, 10.0):
        bq = BatchQP.from_controller(c, rho=rho)
        bq.solve(x0); torch.cuda.synchronize(); t0 = time.perf_counter(); out = bq.solve(x0); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = out["status"]
        if base is None: base = st.clone()
        hard = base == 2
        print(f"N={N} rho={rho}: {dt*1e3:.1f} ms, maxit {(st==2).sum().item()}, of the auto-rho max_iter samples now: solved {(st[hard]==0).sum().item()} infeasible {(st[hard]==1).sum().item()} maxit {(st[hard]==2).sum().item()}, mean iters {bq.last_stats()[0]/x0.shape[1]:.1f}, flag diff vs auto {((st==0)!=(base==0)).sum().item()}")
PY
