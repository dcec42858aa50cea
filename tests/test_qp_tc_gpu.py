"""The tcgen05 ADMM kernel (csrc/qp_admm_tc.cu) through the C ABI: same certified results as the FFMA kernel on the same
states, and the exact oracle's results at BASELINE's tolerances (the reference's QP: lib/mpc.py:318-335)."""
import numpy as np
import pytest

from test_qp_gpu import torch_cuda, _setup, _compare, _assert_same_solution          # noqa: F401

pytestmark = pytest.mark.gpu


def _random_config_states(torch, n, seed):
    from carmpc_b200.grids import config3_axes, materialise_grid
    grid = torch.stack(materialise_grid(config3_axes(), device="cuda")).contiguous()
    idx = torch.randperm(grid.shape[1], generator=torch.Generator().manual_seed(seed))[:n].cuda()
    return grid[:, idx].contiguous()


@pytest.mark.parametrize("env_name,N,n_states", [("RoadOneCarEnv", 10, 40_000), ("RoadOneCarEnv", 20, 60_000),
                                                 ("RoadOneCarEnv", 40, 40_000), ("RoadMultipleCarsEnv", 20, 30_000),
                                                 ("RoadEnv", 40, 30_000), ("RoadOneCarEnv", 80, 30_000), ("RoadEnv", 80, 25_000)])
def test_tensor_kernel_gives_the_ffma_kernels_results(torch_cuda, env_name, N, n_states):
    torch = torch_cuda
    c, bq, oq = _setup(env_name, N)
    assert bq.tensor_mode()["available"]
    assert bq.tensor_mode()["parts"] == (2 if N == 80 else 1)      # horizon 80: one pass per independent chain
    x0 = _random_config_states(torch, n_states, seed=N)
    bq.tensor_mode(0)
    a = {k: v.clone() for k, v in bq.solve(x0, want_u_full=True).items() if torch.is_tensor(v)}
    assert bq.tensor_mode()["samples_last_solve"] == 0
    bq.tensor_mode(2)
    b = bq.solve(x0, want_u_full=True)
    assert bq.tensor_mode()["samples_last_solve"] == n_states, "the first pass did not run on the tcgen05 kernel"
    _assert_same_solution(a, b, oq, x0)
    # undecided states (status 2: 134 of 10^6 on the RoadMultipleCarsEnv grid, none elsewhere) are the same ones
    assert torch.equal(a["status"] == 2, b["status"] == 2) and (b["status"] == 2).sum().item() <= n_states // 2000
    # and against the exact oracle on a subset
    idx = np.arange(0, n_states, n_states // 150)[:150]
    idx = idx[(b["status"].cpu().numpy()[idx] != 2)]
    from carmpc_b200.batch import QPResult
    res = QPResult(u0=b["u0"].cpu().numpy().T[idx], objective=b["objective"].cpu().numpy()[idx],
                   status=b["status"].cpu().numpy()[idx], iters=b["iters"].cpu().numpy()[idx])
    n_band, _, _ = _compare(res, x0.cpu().numpy().T[idx], oq, np.array(c.goal, dtype=float), min_feasible=15)
    assert n_band <= 3


def test_tensor_kernel_modes_and_defaults(torch_cuda):
    torch = torch_cuda
    # default mode 1: horizon 40 (matrices not resident in the FFMA kernel) goes to the tcgen05 kernel, horizon 20 does not
    for N, expect in ((20, 0), (40, 25_000)):
        c, bq, _ = _setup("RoadOneCarEnv", N)
        assert bq.tensor_mode()["mode"] == 1
        x0 = _random_config_states(torch, 25_000, seed=3)
        bq.solve(x0)
        assert bq.tensor_mode()["samples_last_solve"] == expect
        # small batches never use it (narrow FFMA tiles are the right tool there)
        bq.tensor_mode(2)
        bq.solve(x0[:, :5_000].contiguous())
        assert bq.tensor_mode()["samples_last_solve"] == 0
    # cycle counters: the roles' waits add up to less than their totals
    bq.tensor_mode(3)
    bq.solve(x0)
    cyc = bq.tensor_mode()["cycles"]
    assert cyc[8] > 0 and cyc[0] > cyc[1] + cyc[2] and cyc[7] > cyc[3] + cyc[4]


def test_tensor_kernel_with_disturbance_and_warm_start(torch_cuda):
    """Per-sample disturbance (the constant column of the e block) and a warm-started second solve."""
    torch = torch_cuda
    c, bq, oq = _setup("RoadOneCarEnv", 20)
    n = 30_000
    x0 = _random_config_states(torch, n, seed=11)
    g = torch.Generator(device="cpu").manual_seed(5)
    cd = ((torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * 0.2).cuda()
    outs = {}
    for mode in (0, 2):
        bq.tensor_mode(mode)
        outs[mode] = {k: v.clone() for k, v in bq.solve(x0, c=cd).items() if torch.is_tensor(v)}
    sa, sb = outs[0]["status"], outs[2]["status"]
    assert (sa != sb).sum().item() <= 3
    ok = (sa == 0) & (sb == 0)
    assert (outs[0]["u0"] - outs[2]["u0"])[:, ok].abs().max().item() <= 1e-7
    # warm start: the state written by the tcgen05 kernel restarts either kernel in (almost) no iterations
    warm = torch.zeros((n, bq.pq.m + bq.pq.n), dtype=torch.float32, device="cuda")
    bq.tensor_mode(2)
    first = bq.solve(x0, warm=warm, warm_out=True)
    it_cold = bq.last_stats()[0]
    for mode in (0, 2):
        bq.tensor_mode(mode)
        again = bq.solve(x0, warm=warm.clone(), warm_in=True, warm_out=True)
        assert (again["status"] != first["status"]).sum().item() <= 3
        assert bq.last_stats()[0] <= 0.6 * it_cold


@pytest.mark.parametrize("N", [20, 40, 80])
def test_tensor_kernel_disturbance_columns_and_bad_states(torch_cuda, N):
    """MPCOutputFBWithDisturbance has a non-zero Gc (the per-sample disturbance shifts the constraint bounds): the constant
    columns of the tcgen05 kernel (x0 pieces, 1, disturbance pieces - per chain at horizon 80) must reproduce the FFMA
    kernel's results; non-finite states are flagged infeasible by both and do not disturb their tile neighbours."""
    torch = torch_cuda
    from conftest import make_env, make_controller
    from carmpc_b200.batch import BatchQP
    ctl = make_controller(make_env("RoadEnv"), N, cls="MPCOutputFBWithDisturbance", init_state=[20, 0.5, 0, 2])
    bq = BatchQP.from_controller(ctl)
    assert bq.tensor_mode()["parts"] == (2 if N == 80 else 1)
    n = 24_000
    g = torch.Generator(device="cpu").manual_seed(N)
    x_ref = torch.tensor([30.0, 1.5, 0.0, 0.0], dtype=torch.float64)
    spread = torch.tensor([12.0, 1.2, 0.2, 2.0], dtype=torch.float64)
    x0 = (x_ref[:, None] + (torch.rand((4, n), generator=g, dtype=torch.float64) * 2 - 1) * spread[:, None]).cuda().contiguous()
    cd = ((torch.rand(n, generator=g, dtype=torch.float64) - 0.5) * 0.04).cuda()
    x0[0, 5] = float("nan"); x0[1, 130] = float("inf"); x0[3, 4097] = -float("inf"); x0[2, 20_000] = 1e300
    outs = {}
    for mode in (0, 2):
        bq.tensor_mode(mode)
        outs[mode] = {k: v.clone() for k, v in bq.solve(x0, x_ref=x_ref.numpy(), c=cd).items() if torch.is_tensor(v)}
        assert bq.tensor_mode()["samples_last_solve"] == (n if mode else 0)
    sa, sb = outs[0]["status"], outs[2]["status"]
    # (with a disturbance there is no float64 Farkas kernel: barely infeasible states can stay "undecided" (2) in one
    #  kernel and be flagged in the other - never solved in one and infeasible in the other)
    diff = sa != sb
    assert diff.sum().item() <= n // 1000 and bool(((sa[diff] == 2) | (sb[diff] == 2)).all())
    for i in (5, 130, 4097, 20_000):
        assert sb[i].item() == 1
    ok = (sa == 0) & (sb == 0)
    assert ok.sum().item() > n // 10
    assert (outs[0]["u0"] - outs[2]["u0"])[:, ok].abs().max().item() <= 1e-7
    # the disturbance matters on this batch: solving without it changes the answers
    bq.tensor_mode(2)
    plain = bq.solve(x0, x_ref=x_ref.numpy())
    both = ok & (plain["status"] == 0)
    assert (plain["u0"] - outs[2]["u0"])[:, both].abs().max().item() > 1e-4
