"""Whole cold solve (10^6 config-3 states) on the FFMA and the tcgen05 ADMM kernel, alternating, CUDA events:
python tools/tc_compare.py [N] [reps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.grids import config3_axes, materialise_grid
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
x0 = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
bq = BatchQP.from_controller(_controller("RoadOneCarEnv", [29.9, 1.5, 0, 0], N))
for mode in (0, 2):
    bq.tensor_mode(mode); bq.solve(x0)
ms = {0: [], 2: []}
for r in range(reps):
    for mode in (0, 2):
        bq.tensor_mode(mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bq.solve(x0); e1.record(); e1.synchronize()
        ms[mode].append(e0.elapsed_time(e1))
print(f"N={N}: ffma " + " ".join(f"{v:.2f}" for v in ms[0]) + " | tcgen05 " + " ".join(f"{v:.2f}" for v in ms[2]) +
      f" | min {min(ms[0]):.2f} vs {min(ms[2]):.2f} ms")
