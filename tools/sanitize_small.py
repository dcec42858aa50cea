"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck): python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench_qp import _controller
from carmpc_b200.batch import BatchQP, TerminalSetEvaluator, RolloutEvaluator
from carmpc_b200.lib.environments import RoadMultipleCarsEnv
from carmpc_b200.lib.mpc import _C_XYV, _L_OBSERVER

rng = np.random.default_rng(0)
Ab = np.load(os.path.join(ROOT, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
ev = TerminalSetEvaluator(Ab)
for n in (1, 777, 70_001, 1_100_000):          # plain kernel, TMA kernel + ragged tail, auto-tune
    p = np.array([30, 1.5, 0, 0]) + rng.uniform(-1, 1, size=(n, 4)) * np.array([10.0, 2.0, 0.4, 3.0])
    dev = [torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in p.T]
    for mode in (0, 1):
        bits, count = ev.contains_bits(*dev, mode=mode)
    print("membership", n, int(count.item()))
bits, count = ev.contains_grid_bits([np.linspace(5, 55, 20), np.linspace(-3.2, 3.2, 20), np.linspace(-0.42, 0.42, 10), np.linspace(-1.2, 5.2, 10)])
print("grid", int(count.item()))
rv = RolloutEvaluator.from_env(RoadMultipleCarsEnv(), 16)
bits, count, first = rv.contains_bits(*dev, want_first_violation=True)
print("rollout", int(count.item()))
hb, hc = ev.contains_bits_host(*[c.cpu().numpy() for c in dev])
print("host pipeline", hc)
for env, goal, N, B in (("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20, 3000), ("RoadMultipleCarsEnv", None, 10, 500), ("RoadEnv", None, 40, 300),
                        ("RoadOneCarEnv", [29.9, 1.5, 0, 0], 80, 100)):
    c = _controller(env, goal, N)
    bq = BatchQP.from_controller(c)
    x0 = np.array(c.goal, dtype=float) + rng.uniform(-1, 1, size=(B, 4)) * np.array([12.0, 1.4, 0.25, 2.5])
    res = bq.solve_host(x0, want_u_full=True)
    print("qp", env, N, np.bincount(res.status, minlength=3), res.iters.mean())
ofb = _controller("RoadEnv", None, 20)
bl = BatchQP.from_controller(ofb)
x_init = torch.from_numpy(np.ascontiguousarray((np.array([20, 0.5, 0, 2.0]) + rng.uniform(-1, 1, size=(400, 4)) * np.array([8, 1.5, 0.1, 1.0])).T)).cuda()
out = bl.closed_loop(x_init, 12, ofb.A, ofb.B, C=_C_XYV, L=_L_OBSERVER, want_traj=True, want_inputs=True)
print("closed loop", int((out["fail_step"] < 0).sum().item()), out["total_iters"])
# seeded maps: anchors' multiplier maps / Farkas certificates, followers, certificate filter, map entry point
from carmpc_b200.grids import lattice_seeds, materialise_grid
for env, goal, N in (("RoadOneCarEnv", [29.9, 1.5, 0, 0], 20), ("RoadMultipleCarsEnv", None, 10), ("RoadEnv", None, 40)):
    c = _controller(env, goal, N)
    bq = BatchQP.from_controller(c)
    g = np.array(c.goal, dtype=float)
    axes = [np.linspace(g[0] - 20.0, g[0] + 0.5, 14), np.linspace(-3.0, 3.0, 26), np.linspace(-0.3, 0.3, 3), np.linspace(-1.0, 4.0, 3)]
    x0 = torch.stack(materialise_grid(axes, device="cuda")).contiguous()
    seed = torch.from_numpy(lattice_seeds([len(a) for a in axes], block=(3, 8, 1, 1))).cuda()
    out = bq.solve(x0, seed=seed, want_u_full=True)
    res = bq.solve_map_host(axes, block=(2, 5, 2, 1), pinned=True)
    print("seeded", env, N, out["seeded"], res.seeded, np.bincount(res.status, minlength=3), bq.polish_stats()["infeasible_by_anchor_certificate"])
# ragged implicit grids
for dims in ((3, 5, 7, 11), (1, 1, 1, 1), (2, 1, 129, 1), (13, 2, 3, 257)):
    ax = [np.linspace(20.0, 40.0, d) if d > 1 else np.array([30.0]) for d in dims]
    ax[1] = ax[1] * 0.05; ax[2] = ax[2] * 0.0; ax[3] = ax[3] * 0.02
    bits, count = ev.contains_grid_bits(ax)
    print("grid", dims, int(count.item()))
