"""Closed loop for profiling: python tools/prof_cl.py [runs] [steps] [reps]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP
from carmpc_b200.lib.mpc import _C_XYV as C_OUT, _L_OBSERVER as L_OBS
R = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
WS = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ofb = _controller("RoadEnv", None, 20)
bl = BatchQP.from_controller(ofb)
g = torch.Generator(device="cpu").manual_seed(0)
lo = torch.tensor([0.0, -2.5, -0.2, 0.0], dtype=torch.float64); hi = torch.tensor([10.0, 2.5, 0.2, 3.0], dtype=torch.float64)
x_init = (lo[:, None] + (hi - lo)[:, None] * torch.rand((4, R), generator=g, dtype=torch.float64)).cuda().contiguous()
bl.closed_loop(x_init, 3, ofb.A, ofb.B, C=C_OUT, L=L_OBS, warm_start=WS)
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = bl.closed_loop(x_init, T, ofb.A, ofb.B, C=C_OUT, L=L_OBS, warm_start=WS)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"warm_start {WS} runs {R} steps {T}: {dt*1e3:.1f} ms ({dt/T*1e6:.0f} us/step), alive {(o['fail_step'] < 0).float().mean().item():.3f}, iters {o['total_iters']}")
