"""Diagnostic: which GPU results fail the reference-KKT check, and by how much (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from kkt_check import ReferenceQP
from test_qp_gpu import _setup, _config_states

cases = [("RoadOneCarEnv", 20, 10000), ("RoadEnv", 20, 10000), ("RoadMultipleCarsEnv", 20, 10000), ("RoadOneCarEnv", 10, 10000),
         ("RoadOneCarEnv", 40, 4000), ("RoadOneCarEnv", 80, 2000)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:] or str(c[1]) in sys.argv[1:]]
for env_name, N, n_states in cases:
    ref = ReferenceQP(env_name, N)
    c, bq, oq = _setup(env_name, N)
    x0 = _config_states(ref, n_states, seed=N)
    res = bq.solve_host(x0, want_u_full=True)
    ok = res.status == 0
    primal, stat, nact = ref.kkt(res.u_full[ok], x0[ok])
    idx = np.flatnonzero(ok)
    badp = primal > 1e-8
    bads = stat > 1e-6
    print(f"== {env_name} N={N}: solved {ok.sum()}, status2 {(res.status == 2).sum()}, primal>1e-8: {badp.sum()} (max {primal.max():.3e}), "
          f"stat>1e-6: {bads.sum()} (max {stat.max():.3e}), polish {bq.polish_stats()}")
    off = np.flatnonzero(badp | bads)[:12]
    if len(off):
        from oracle import carmpc_oracle as orc
        ue, obje, ste, pol, slack = orc.qp_solve_exact(oq, x0[idx[off]], ref.goal)
        for k, j in enumerate(off):
            i = idx[j]
            print(f"   sample {i} x0={np.array2string(x0[i], precision=4)} iters={res.iters[i]} primal={primal[j]:.3e} stat={stat[j]:.3e} nact={nact[j]} "
                  f"oracle: status={ste[k]} slack={slack[k]:.3e} polished={pol[k]} |du|={np.abs(res.u_full[i] - ue[k]).max():.3e} "
                  f"dobj={res.objective[i] - obje[k]:.3e}")
