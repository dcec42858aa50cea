"""Parity of the CUDA terminal-set kernels against the oracle, through the C ABI (needs a B200).

Bit-exact: the C oracle evaluates the same fma chain in IEEE float64, so bitsets are compared with ==.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, FIXTURES, K_STAR, golden, make_env

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "the GPU tests need a CUDA device"
    return torch


def _dev(torch, *arrays):
    return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda() for a in arrays]


def _bits_np(bits_tensor):
    return bits_tensor.cpu().numpy().view(np.uint32)


def test_config1_grid_known_answer(torch_cuda):
    """lib/terminal_set.py:96-113 on RoadOneCarEnv goal (29.9, 1.5, 0, 0): 434 members, [49, 70, 84, 91, 91, 49]."""
    from carmpc_b200.batch import TerminalSetEvaluator, unpack_bits
    from carmpc_b200.lib.terminal_set import grid_points, grid_membership
    g = golden("grid_config1.npz")
    env = make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0])
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    x, y, psi, v = grid_points(env)
    ev = TerminalSetEvaluator(Ab)
    want = g["member"].reshape(-1)
    for mode in (0, 1):
        bits, count = ev.contains_bits(*_dev(torch_cuda, x, y, psi, v), mode=mode)
        got = unpack_bits(_bits_np(bits), len(x))
        np.testing.assert_array_equal(got, want)
        assert int(count.item()) == 434
        assert list(got.reshape(6, -1).sum(1)) == [49, 70, 84, 91, 91, 49]
    # the drop-in entry point the reference's visualise_set would call
    pts, member = grid_membership(Ab[:, :4], Ab[:, 4], env)
    np.testing.assert_array_equal(member, want)
    assert pts.shape == (60000, 4)
    # implicit grid: axes (v, y, x) with psi a single-point axis
    xs = np.linspace(4.9, 54.9, 100)
    ys = np.linspace(-23.5, 26.5, 100)
    bits, count = ev.contains_grid_bits([np.arange(6.0), ys, xs, [0.0]], axis_to_state=(3, 1, 0, 2))
    np.testing.assert_array_equal(unpack_bits(_bits_np(bits), 60000), want)
    assert int(count.item()) == 434


@pytest.mark.parametrize("file", sorted(FIXTURES))
@pytest.mark.parametrize("n", [1, 31, 32, 33, 511, 513, 100_003, 1_000_000])
def test_membership_bit_exact_vs_oracle(torch_cuda, file, n):
    from carmpc_b200.batch import TerminalSetEvaluator
    from oracle import c_oracle
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", file))
    rng = np.random.default_rng(n)
    goal = np.array(FIXTURES[file][1], dtype=float)
    p = goal + rng.uniform(-1, 1, size=(n, 4)) * np.array([12.0, 2.5, 0.45, 3.2])
    p[: n // 4] = goal + rng.uniform(-1, 1, size=(n // 4, 4)) * np.array([30.0, 6.0, 1.0, 8.0])
    want_bits, want_cnt = c_oracle.membership_bits(Ab, *p.T)
    ev = TerminalSetEvaluator(Ab)
    x, y, psi, v = _dev(torch_cuda, *p.T)
    for mode in (0, 1):
        bits, count = ev.contains_bits(x, y, psi, v, mode=mode)
        np.testing.assert_array_equal(_bits_np(bits), want_bits)
        assert int(count.item()) == want_cnt


def test_membership_edge_cases(torch_cuda):
    from carmpc_b200.batch import TerminalSetEvaluator
    from carmpc_b200._capi import CarmpcError
    from oracle import c_oracle
    torch = torch_cuda
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    ev = TerminalSetEvaluator(Ab)
    # empty input
    e = torch.empty(0, dtype=torch.float64, device="cuda")
    bits, count = ev.contains_bits(e, e, e, e)
    assert bits.numel() == 0 and int(count.item()) == 0
    # unaligned views (odd element offset -> scalar-load path), ragged tail
    rng = np.random.default_rng(5)
    p = np.array([30, 1.5, 0, 0]) + rng.uniform(-1, 1, size=(4099, 4)) * np.array([10.0, 2.0, 0.4, 3.0])
    big = [torch.from_numpy(np.concatenate(([0.0], c))).cuda() for c in p.T]
    views = [b[1:] for b in big]
    want_bits, want_cnt = c_oracle.membership_bits(Ab, *p.T)
    bits, count = ev.contains_bits(*views, mode=1)
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    # exact ties on a facet: v = 5 with '<='
    tie = np.array([[30.0, 1.5, 0.0, 5.0], [30.0, 1.5, 0.0, np.nextafter(5.0, 6.0)]])
    bits, count = ev.contains_bits(*_dev(torch, *tie.T), mode=1)
    want_bits, _ = c_oracle.membership_bits(Ab, *tie.T)
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    # empty polytope: everything is inside
    ev0 = TerminalSetEvaluator(np.zeros((0, 5)))
    bits, count = ev0.contains_bits(*_dev(torch, *p.T))
    assert int(count.item()) == len(p)
    # argument errors surface as exceptions with the library's message
    with pytest.raises(ValueError):
        ev.contains_bits(views[0].float(), views[1], views[2], views[3])
    with pytest.raises(CarmpcError):
        ev.contains_bits(*views, mode=7)
    # NaN coordinates are never members (NaN <= b is false)
    nanp = np.array([[np.nan, 1.5, 0.0, 1.0], [30.0, 1.5, 0.0, 0.0]])
    bits, count = ev.contains_bits(*_dev(torch, *nanp.T), mode=1)
    assert int(count.item()) == 1 and int(_bits_np(bits)[0]) == 2


@pytest.mark.parametrize("file", sorted(FIXTURES))
def test_rollout_bit_exact_vs_oracle_and_equals_hrep(torch_cuda, file):
    from carmpc_b200.batch import RolloutEvaluator, TerminalSetEvaluator, unpack_bits
    from oracle import c_oracle, carmpc_oracle as orc
    env_name, goal = FIXTURES[file]
    env = make_env(env_name, goal)
    k = K_STAR[env_name]
    n = 300_007
    rng = np.random.default_rng(7)
    g = np.array(goal, dtype=float)
    p = g + rng.uniform(-1, 1, size=(n, 4)) * np.array([12.0, 2.5, 0.45, 3.0])
    for every in (False, True):
        ev = RolloutEvaluator.from_env(env, k, input_every_step=every)
        want_bits, want_first, want_cnt = c_oracle.rollout_bits(ev.A_k, ev.A_con, ev.b_con, ev.A_in, ev.b_in, g, k,
                                                                int(every), *p.T)
        bits, count, first = ev.contains_bits(*_dev(torch_cuda, *p.T), want_first_violation=True)
        np.testing.assert_array_equal(_bits_np(bits), want_bits)
        np.testing.assert_array_equal(first.cpu().numpy(), want_first)
        assert int(count.item()) == want_cnt
        # without the per-sample step output the float32-screened path runs (expanded rows a_r A_k^t, float64 rollout
        # for the samples the screen cannot decide): same bits
        bits, count = ev.contains_bits(*_dev(torch_cuda, *p.T))
        np.testing.assert_array_equal(_bits_np(bits), want_bits)
        assert int(count.item()) == want_cnt
    # the oracle's setup and the product's are the same numbers
    Ak, K, Ac, bc, Ai, bi = orc.rollout_setup(env_name, goal)
    ev = RolloutEvaluator.from_env(env, k)
    np.testing.assert_allclose(ev.A_k, Ak, atol=1e-13)
    np.testing.assert_allclose(ev.A_con, Ac, atol=1e-13)
    np.testing.assert_allclose(ev.b_con, bc, atol=1e-12)
    np.testing.assert_allclose(ev.A_in, Ai, atol=1e-12)
    # rollout form == shipped H-rep outside the 1e-6 boundary band, band enumerated
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", file))
    bits_r, _ = ev.contains_bits(*_dev(torch_cuda, *p.T))
    bits_h, _ = TerminalSetEvaluator(Ab).contains_bits(*_dev(torch_cuda, *p.T))
    diff = unpack_bits(_bits_np(bits_r), n) != unpack_bits(_bits_np(bits_h), n)
    _, margin_h = orc.membership(Ab, *p.T)
    _, _, margin_r = orc.rollout_membership(env_name, goal, k, *p.T)
    band = (np.abs(margin_h) <= 1e-6) | (np.abs(margin_r) <= 1e-6)
    assert not (diff & ~band).any(), f"{(diff & ~band).sum()} disagreements outside the boundary band"


@pytest.mark.parametrize("env_name", ["RoadOneCarEnv", "RoadMultipleCarsEnv", "RoadEnv"])
def test_screened_rollout_on_boundary_samples(torch_cuda, env_name):
    """Samples placed on (and a few ulps around) facets of the rollout set, ties included, plus non-finite and huge
    coordinates: the screened path must return the float64 step-by-step decision of the oracle for every one."""
    from carmpc_b200.batch import RolloutEvaluator
    from oracle import c_oracle, carmpc_oracle as orc
    env = make_env(env_name)
    goal = np.array(env.goal, dtype=float)
    k = K_STAR[env_name]
    ev = RolloutEvaluator.from_env(env, k)
    Ak, K, Ac, bc, Ai, bi = orc.rollout_setup(env_name, list(goal))
    rng = np.random.default_rng(23)
    rows, M = [], np.eye(4)
    for t in range(k + 1):
        for a, b in zip(Ac, bc):
            rows.append((a @ M, b))
        if t == 0:
            for a, b in zip(Ai, bi):
                rows.append((a @ M, b))
        M = Ak @ M
    pts = []
    for g, b in rows:
        if not np.isfinite(b) or np.linalg.norm(g) < 1e-9:
            continue
        for _ in range(40):
            e = rng.uniform(-1, 1, 4) * np.array([10.0, 2.0, 0.4, 3.0])
            e = e + g * (b - g @ e) / (g @ g)                       # projected onto the facet g e = b
            for scale in (0.0, 1e-15, -1e-15, 3e-9, -3e-9, 1e-6, -1e-6):
                pts.append(goal + e + scale * g / np.linalg.norm(g))
    pts = np.array(pts)
    special = np.array([[np.nan, 1.5, 0, 0], [30, np.inf, 0, 0], [30, 1.5, -np.inf, 0], [1e300, 1.5, 0, 0], [-1e300, 1e300, 0, 0],
                        [1e38, 0, 0, 0], list(goal), [30, 1.5, 0, 1e-310]])
    pts = np.vstack((pts, special, pts[:37]))
    cols = [np.ascontiguousarray(c) for c in pts.T]
    want_bits, _, want_cnt = c_oracle.rollout_bits(ev.A_k, ev.A_con, ev.b_con, ev.A_in, ev.b_in, goal, k, 0, *cols)
    bits, count = ev.contains_bits(*_dev(torch_cuda, *cols))
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    assert int(count.item()) == want_cnt and 0 < want_cnt < len(pts)
    # large aligned batch through the bulk-async path: tile the boundary samples past 64 chunks
    reps = (70 * 1024) // len(pts) + 1
    big = [np.ascontiguousarray(np.tile(c, reps)) for c in cols]
    want_bits, _, want_cnt = c_oracle.rollout_bits(ev.A_k, ev.A_con, ev.b_con, ev.A_in, ev.b_in, goal, k, 0, *big)
    bits, count = ev.contains_bits(*_dev(torch_cuda, *big))
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    assert int(count.item()) == want_cnt


def test_host_pipeline_multi_chunk(torch_cuda):
    """Host-buffer entry points: > 1 pipeline chunk (8 Mi samples), ragged tail, same bits as the device path."""
    from carmpc_b200.batch import TerminalSetEvaluator, RolloutEvaluator
    from oracle import c_oracle
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    n = (1 << 23) * 2 + 12_345
    rng = np.random.default_rng(11)
    p = np.array([30, 1.5, 0, 0]) + rng.uniform(-1, 1, size=(n, 4)) * np.array([10.0, 2.0, 0.4, 3.0])
    cols = [np.ascontiguousarray(c) for c in p.T]
    want_bits, want_cnt = c_oracle.membership_bits(Ab, *cols)
    ev = TerminalSetEvaluator(Ab)
    for mode in (0, 1):
        bits, cnt = ev.contains_bits_host(*cols, mode=mode)
        np.testing.assert_array_equal(bits, want_bits)
        assert cnt == want_cnt
    env = make_env("RoadMultipleCarsEnv")
    rv = RolloutEvaluator.from_env(env, 16)
    m = 3_000_001
    want_bits, want_first, want_cnt = c_oracle.rollout_bits(rv.A_k, rv.A_con, rv.b_con, rv.A_in, rv.b_in, rv.goal, 16, 0,
                                                            *[c[:m] for c in cols])
    bits, cnt, first = rv.contains_bits_host(*[c[:m] for c in cols], want_first_violation=True)
    np.testing.assert_array_equal(bits, want_bits)
    np.testing.assert_array_equal(first, want_first)
    assert cnt == want_cnt
    bits, cnt = rv.contains_bits_host(*[c[:m] for c in cols])[:2]                  # screened path
    np.testing.assert_array_equal(bits, want_bits)
    assert cnt == want_cnt


def test_config2_full_grid_properties(torch_cuda):
    """BASELINE config 2: RoadMultipleCarsEnv on the 100^4 = 10^8 grid.  The oracle cannot scan 10^8 points in
    seconds, so: (i) the implicit-grid kernel and the explicit SoA kernel give the same bitset and count,
    (ii) a 2 % strided sample is bit-exact against the C oracle, (iii) the count is the sum of shard counts."""
    torch = torch_cuda
    from carmpc_b200.batch import TerminalSetEvaluator, unpack_bits
    from carmpc_b200.grids import config2_axes, materialise_grid
    from oracle import c_oracle
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    ev = TerminalSetEvaluator(Ab)
    axes = config2_axes()
    x, y, psi, v = materialise_grid(axes, device="cuda")
    n = x.numel()
    assert n == 10 ** 8
    bits, count = ev.contains_bits(x, y, psi, v, mode=1)
    bits0, count0 = ev.contains_bits(x, y, psi, v, mode=0)
    assert torch.equal(bits, bits0) and int(count.item()) == int(count0.item())
    gbits, gcount = ev.contains_grid_bits(axes)
    assert torch.equal(bits, gbits) and int(count.item()) == int(gcount.item())
    frac = int(count.item()) / n
    assert 0.001 < frac < 0.2
    # strided sample against the oracle
    idx = torch.arange(0, n, 47, device="cuda")
    sx, sy, sp, sv = [t[idx].cpu().numpy() for t in (x, y, psi, v)]
    want_bits, _ = c_oracle.membership_bits(Ab, sx, sy, sp, sv)
    got = unpack_bits(_bits_np(bits), n)[idx.cpu().numpy()]
    np.testing.assert_array_equal(got, c_oracle.unpack_bits(want_bits, len(sx)))
    # shard additivity (the multi-GPU partition: contiguous ranges, multiples of 32)
    total = 0
    for r in range(8):
        lo, hi = r * (n // 8), (r + 1) * (n // 8)
        b, c = ev.contains_bits(x[lo:hi], y[lo:hi], psi[lo:hi], v[lo:hi])
        assert torch.equal(b, bits[lo // 32: hi // 32])
        total += int(c.item())
    assert total == int(count.item())
    # the whole grid against the C oracle (multi-threaded: 10^8 points in about a second), H-rep and rollout form.
    # Known answers (reproduced by the oracle on the CPU): 3,028,578 members of the shipped H-rep, 3,026,913 of its
    # rollout form with k* = 16; the 1,665 differences are exact ties of the H-rep (margin 0) that the rollout form
    # misses by one rounding of p - goal (margin -7.8e-16): boundary-band samples, enumerated here.
    from carmpc_b200.batch import RolloutEvaluator
    host = [t.cpu().numpy() for t in (x, y, psi, v)]
    want_bits, want_cnt = c_oracle.membership_bits(Ab, *host)
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    assert int(count.item()) == want_cnt == 3_028_578
    rv = RolloutEvaluator.from_env(make_env("RoadMultipleCarsEnv"), 16)
    rbits, rcount = rv.contains_bits(x, y, psi, v)
    want_rbits, _, want_rcnt = c_oracle.rollout_bits(rv.A_k, rv.A_con, rv.b_con, rv.A_in, rv.b_in, rv.goal, 16, 0, *host)
    np.testing.assert_array_equal(_bits_np(rbits), want_rbits)
    assert int(rcount.item()) == want_rcnt == 3_026_913
    differ = np.flatnonzero(unpack_bits(_bits_np(bits), n) != unpack_bits(_bits_np(rbits), n))
    assert len(differ) == 1665
    from oracle import carmpc_oracle as orc
    pts = [h[differ] for h in host]
    _, margin_h = orc.membership(Ab, *pts)
    _, _, margin_r = orc.rollout_membership("RoadMultipleCarsEnv", [30, 1.5, 0, 0], 16, *pts)
    assert np.abs(margin_h).max() <= 1e-6 and np.abs(margin_r).max() <= 1e-6


def test_profile_guided_row_order_never_changes_results(torch_cuda):
    """The library re-orders the H-rep rows after a pilot over the samples (rows that reject most first).  The
    conjunction over rows is order-independent: bitsets stay bit-exact against the oracle, before and after tuning,
    on the tuning distribution and on a different one, and for a set with more than 64 rows (tuning skipped)."""
    from carmpc_b200.batch import TerminalSetEvaluator
    from oracle import c_oracle
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    rng = np.random.default_rng(21)
    n = 1_200_000
    goal = np.array([30, 1.5, 0, 0.0])
    p1 = goal + rng.uniform(-1, 1, size=(n, 4)) * np.array([25.0, 5.0, 0.8, 6.0])       # mostly outside
    p2 = goal + rng.uniform(-1, 1, size=(n, 4)) * np.array([3.0, 0.5, 0.1, 1.0])        # mostly inside
    ev = TerminalSetEvaluator(Ab)
    for p in (p2[:5000], p1, p2, p1[:777]):        # small call (no tuning yet), large call (tunes), other distribution
        want_bits, want_cnt = c_oracle.membership_bits(Ab, *p.T)
        for mode in (1, 0):
            bits, count = ev.contains_bits(*_dev(torch_cuda, *p.T), mode=mode)
            np.testing.assert_array_equal(_bits_np(bits), want_bits)
            assert int(count.item()) == want_cnt
    ev.tune(*_dev(torch_cuda, *p2.T))               # explicit re-tune on the other distribution
    want_bits, want_cnt = c_oracle.membership_bits(Ab, *p1.T)
    bits, count = ev.contains_bits(*_dev(torch_cuda, *p1.T))
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    # > 64 rows: tuning is skipped, results still exact
    big = np.vstack([Ab, Ab + np.array([0, 0, 0, 0, 0.5])])
    assert len(big) > 64
    evb = TerminalSetEvaluator(big)
    want_bits, want_cnt = c_oracle.membership_bits(big, *p1.T)
    bits, count = evb.contains_bits(*_dev(torch_cuda, *p1.T))
    np.testing.assert_array_equal(_bits_np(bits), want_bits)
    assert int(count.item()) == want_cnt


@pytest.mark.parametrize("dims", [(3, 5, 7, 11), (1, 1, 1, 1), (2, 1, 129, 1), (1, 37, 1, 5), (13, 2, 3, 257)])
def test_implicit_grid_equals_explicit_points_on_ragged_grids(torch_cuda, dims):
    """carmpc_membership_grid generates the coordinates in-kernel (one index decomposition per thread, then increments
    with carry): odd axis lengths, single-point axes and sizes that are no multiple of a tile must give the bits of the
    explicit SoA call."""
    torch = torch_cuda
    from carmpc_b200.batch import TerminalSetEvaluator, unpack_bits
    from carmpc_b200.grids import materialise_grid
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    ev = TerminalSetEvaluator(Ab)
    lo, hi = np.array([20.0, -1.0, -0.2, -0.5]), np.array([40.0, 3.5, 0.2, 1.5])
    axes = [np.linspace(lo[k], hi[k], d) if d > 1 else np.array([0.5 * (lo[k] + hi[k])]) for k, d in enumerate(dims)]
    n = int(np.prod(dims))
    x, y, psi, v = materialise_grid(axes, device="cuda")
    want_bits, want_count = ev.contains_bits(x, y, psi, v, mode=0)
    want = unpack_bits(_bits_np(want_bits), n)
    bits, count = ev.contains_grid_bits(axes)
    np.testing.assert_array_equal(unpack_bits(_bits_np(bits), n), want)
    assert int(count.item()) == int(want_count.item()) == int(want.sum())
    # the same grid with the axes given in another order
    perm = (2, 0, 3, 1)
    bits2, count2 = ev.contains_grid_bits([axes[k] for k in perm], axis_to_state=perm)
    got2 = unpack_bits(_bits_np(bits2), n).reshape([dims[k] for k in perm])
    np.testing.assert_array_equal(np.transpose(got2, np.argsort(perm)).ravel(), want)


def test_reduced_rollout_screen_gives_the_full_screens_bits(torch_cuda):
    """The float32 screen keeps only the irredundant expanded rows (42 of 140 for RoadMultipleCarsEnv); decisions are those
    of the full screen and of the exact kernel, and certificates that do not hold are refused."""
    import ctypes
    from carmpc_b200 import _capi
    from carmpc_b200.batch import RolloutEvaluator, CarmpcError
    env = make_env("RoadMultipleCarsEnv", [30, 1.5, 0, 0])
    k = K_STAR["RoadMultipleCarsEnv"]
    red, full = RolloutEvaluator.from_env(env, k), RolloutEvaluator.from_env(env, k)
    full2 = RolloutEvaluator(full.A_k, full.A_con, full.b_con, full.A_in, full.b_in, full.goal, k, reduce_screen=False)
    assert red.expanded_rows == 140 and red.screen_rows <= 48 and full2.screen_rows == 0
    rng = np.random.default_rng(3)
    n = 2_000_000 + 37
    g = np.array(env.goal, dtype=float)
    p = g + rng.uniform(-1, 1, size=(n, 4)) * np.array([8.0, 2.5, 0.45, 3.0])
    dev = _dev(torch_cuda, *p.T)
    b_red, c_red = red.contains_bits(*dev)
    b_full, c_full = full2.contains_bits(*dev)
    b_exact, c_exact, _ = full2.contains_bits(*dev, want_first_violation=True)
    assert int(c_red.item()) == int(c_full.item()) == int(c_exact.item()) > 1000
    np.testing.assert_array_equal(_bits_np(b_red), _bits_np(b_full))
    np.testing.assert_array_equal(_bits_np(b_red), _bits_np(b_exact))
    # a certificate that does not reproduce the dropped row is refused and leaves the handle as it was
    rows = full2.expanded()
    kept, idx, w = RolloutEvaluator.irredundant_rows(rows)
    bad = w.copy()
    bad[bad > 0] *= 0.5
    lib = _capi.load()
    rc = lib.carmpc_rollout_reduce_screen(full2._h, kept.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(kept),
                                          idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _capi.ptr(bad), idx.shape[1])
    assert rc != 0
    b_again, c_again = full2.contains_bits(*dev)
    np.testing.assert_array_equal(_bits_np(b_again), _bits_np(b_full))
