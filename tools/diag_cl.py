"""Diagnostic: closed-loop fixture mismatches (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_qp_gpu import _setup
from carmpc_b200.lib.mpc import _C_XYV, _L_OBSERVER
from oracle import carmpc_oracle as orc
fx = np.load(os.path.join(ROOT, "tests", "golden", "closed_loop_config4.npz"))
c, bq, oq = _setup("RoadEnv", 20)
g = np.array(c.goal, float)
x_init = fx["x_init"]; steps = int(fx["steps"]); stride = int(fx["stride"])
for tag, fb in (("ofb", True), ("sfb", False)):
    out = bq.closed_loop(torch.from_numpy(np.ascontiguousarray(x_init.T)).cuda(), steps, c.A, c.B, C=_C_XYV if fb else None,
                         L=_L_OBSERVER if fb else None, want_traj=True, want_inputs=True)
    fail = out["fail_step"].cpu().numpy(); want = fx[f"fail_{tag}"]
    traj = out["traj"].cpu().numpy().transpose(0, 2, 1)
    bad = np.flatnonzero(fail != want)
    print(tag, "mismatches", len(bad), "of", len(want))
    for r in bad[:10]:
        k = min(x for x in (fail[r], want[r]) if x >= 0)
        print(f"  run {r}: gpu fail {fail[r]} oracle fail {want[r]}; x_init {x_init[r]}")
        # state fed to the QP at step k (state feedback: the plant state after the step-k plant move)
        xs = traj[k, r]
        if not fb:
            feas, slack = orc.qp_feasible_lp(oq, xs[None, :])
            print(f"     state at step {k}: {xs} oracle LP slack {slack[0]:.3e}")
        near = [s for s in range(stride - 1, steps, stride) if s < k]
        if near:
            s = near[-1]
            print(f"     traj diff at step {s}: {np.abs(traj[s, r] - fx[f'traj_{tag}'][s // stride, r]).max():.3e}")
