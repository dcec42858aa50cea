"""Sharded scans with the all-gather fused into the kernel (carmpc_shard_*), on ONE device: the ranks of a
group live in one process, each with its own window and stream; the peer stores and the flag protocol are exactly the
ones that run between GPUs (the bench exercises the CUDA-IPC form at 2/4/8 GPUs and asserts the same equalities)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


_STREAMS = []


def _streams(torch, k):
    """One fixed set of streams for every test of this module.  The ranks of a group run as streams of ONE process here
    and their exchange kernels wait for each other, so no two of them may share a hardware queue (between processes - one
    rank per GPU - the question does not arise).  Streams from torch's pool are assigned to queues in creation order;
    re-using the same eight keeps the assignment fixed instead of walking through the pool test by test."""
    while len(_STREAMS) < 8:
        _STREAMS.append(torch.cuda.Stream())
    return _STREAMS[:k]


def _samples(torch, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lo = torch.tensor([5.0, -3.2, -0.42, -1.2], dtype=torch.float64, device="cuda")
    hi = torch.tensor([55.0, 3.2, 0.42, 5.2], dtype=torch.float64, device="cuda")
    u = torch.rand((4, n), generator=g, dtype=torch.float64, device="cuda")
    return [(lo[k] + (hi[k] - lo[k]) * u[k]).contiguous() for k in range(4)]


@pytest.mark.parametrize("layout", ["cyclic", "contiguous"])
@pytest.mark.parametrize("world,n", [(2, 3_000_077), (3, 1_500_000), (8, 9_000_001), (4, 5000), (2, 0), (3, 2048), (4, 1_000_000)])
def test_fused_sharded_membership_equals_single_scan(world, n, layout):
    import torch
    from carmpc_b200.batch import TerminalSetEvaluator
    from carmpc_b200.sharding import PeerWindow, contains_bits_sharded, wait_sharded
    ev = TerminalSetEvaluator(np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy")))
    x, y, psi, v = _samples(torch, n, seed=world)
    words = (n + 31) // 32
    if n:
        want_bits, want_count = ev.contains_bits(x, y, psi, v)
        want_bits, want_count = want_bits.clone(), int(want_count.item())
    else:
        want_bits, want_count = torch.empty(0, dtype=torch.int32, device="cuda"), 0
    wins = PeerWindow.local_group(n, world, layout=layout)
    streams = _streams(torch, world)
    totals = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    # the ranks' index sets partition the sample set
    idx = [w.local_index("cuda") for w in wins]
    allidx = torch.cat(idx)
    assert allidx.numel() == n and (n == 0 or torch.equal(torch.sort(allidx).values, torch.arange(n, device="cuda")))
    local = [[t[i].contiguous() for t in (x, y, psi, v)] for i in idx]
    torch.cuda.synchronize()
    for step in range(3):                                   # three steps: both buffer slots, flags keep counting
        for r, w in enumerate(wins):                        # every rank's scan + publish first ...
            with torch.cuda.stream(streams[r]):
                contains_bits_sharded(ev, w, *local[r], total=totals[r], defer_wait=True)
        for r, w in enumerate(wins):                        # ... then every rank's wait (one process drives all ranks here)
            with torch.cuda.stream(streams[r]):
                wait_sharded(w, totals[r])
        torch.cuda.synchronize()
        for r, w in enumerate(wins):
            w.check()
            assert int(totals[r].item()) == want_count
            got = w.result_bits()
            assert got.numel() == words and torch.equal(got, want_bits), f"rank {r} step {step}: gathered bitset differs"


def test_fused_sharded_rollout_equals_single_scan():
    import torch
    from carmpc_b200.batch import RolloutEvaluator
    from carmpc_b200.lib.environments import RoadMultipleCarsEnv
    from carmpc_b200.sharding import PeerWindow, contains_bits_sharded, wait_sharded
    rv = RolloutEvaluator.from_env(RoadMultipleCarsEnv(), 16)
    n, world = 2_100_000, 3
    x, y, psi, v = _samples(torch, n, seed=5)
    want_bits, want_count = rv.contains_bits(x, y, psi, v)
    want_bits, want_count = want_bits.clone(), int(want_count.item())
    wins = PeerWindow.local_group(n, world)
    streams = _streams(torch, world)
    totals = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    local = [[t[w.local_index("cuda")].contiguous() for t in (x, y, psi, v)] for w in wins]
    torch.cuda.synchronize()
    for r, w in enumerate(wins):
        with torch.cuda.stream(streams[r]):
            contains_bits_sharded(rv, w, *local[r], total=totals[r], defer_wait=True)
    for r, w in enumerate(wins):
        with torch.cuda.stream(streams[r]):
            wait_sharded(w, totals[r])
    torch.cuda.synchronize()
    for r, w in enumerate(wins):
        w.check()
        assert int(totals[r].item()) == want_count and torch.equal(w.result_bits(), want_bits)


def test_shard_argument_errors():
    import torch
    from carmpc_b200._capi import CarmpcError
    from carmpc_b200.batch import TerminalSetEvaluator
    from carmpc_b200.sharding import PeerWindow, contains_bits_sharded, wait_sharded
    ev = TerminalSetEvaluator(np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy")))
    x, y, psi, v = _samples(torch, 4096, seed=1)
    w = PeerWindow(4096, rank=0, world=2, _connect=False, layout="contiguous")          # never connected to rank 1
    with pytest.raises(CarmpcError, match="not connected"):
        contains_bits_sharded(ev, w, x[:2048], y[:2048], psi[:2048], v[:2048])
    with pytest.raises(ValueError):
        contains_bits_sharded(ev, w, x, y, psi, v)                 # not this rank's shard
    with pytest.raises(CarmpcError):
        PeerWindow(100, rank=0, world=9, _connect=False)


@pytest.mark.parametrize("geometry", [(256, 3, 1), (128, 3, 1), (128, 4, 1), (128, 2, 2), (128, 3, 2), (256, 2, 1), (128, 2, 1), (256, 2, 2)])
def test_every_staging_geometry_gives_the_same_bits(geometry):
    import torch
    from carmpc_b200._capi import CarmpcError
    from carmpc_b200.batch import TerminalSetEvaluator
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    x, y, psi, v = _samples(torch, 2_000_333, seed=9)
    ref = TerminalSetEvaluator(Ab)
    want_bits, want_count = ref.contains_bits(x, y, psi, v, mode=0)
    ev = TerminalSetEvaluator(Ab)
    ev.set_staging(*geometry)
    bits, count = ev.contains_bits(x, y, psi, v, mode=1)
    assert torch.equal(bits, want_bits) and int(count.item()) == int(want_count.item())
    with pytest.raises(CarmpcError, match="not built"):
        ev.set_staging(96, 3, 1)


def test_collective_step_replays_from_a_cuda_graph():
    """The launch arguments of a collective step never change (the step counter and the buffer slot live on the device),
    so a step captured once in a CUDA graph can be replayed: two ranks on two streams, three replays each."""
    import torch
    from carmpc_b200.batch import TerminalSetEvaluator
    from carmpc_b200.sharding import PeerWindow, contains_bits_sharded, wait_sharded
    Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
    n, world = 2_500_123, 2
    x, y, psi, v = _samples(torch, n, seed=21)
    evs = [TerminalSetEvaluator(Ab) for _ in range(world)]
    want_bits, want_count = evs[0].contains_bits(x, y, psi, v)
    want_bits, want_count = want_bits.clone(), int(want_count.item())
    wins = PeerWindow.local_group(n, world)
    local = [[t[w.local_index("cuda")].contiguous() for t in (x, y, psi, v)] for w in wins]
    totals = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    streams = _streams(torch, world)
    for r in range(world):                                   # eager warm-up step (row-order tuning, attributes)
        with torch.cuda.stream(streams[r]):
            contains_bits_sharded(evs[r], wins[r], *local[r], total=totals[r], defer_wait=True)
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            wait_sharded(wins[r], totals[r])
    torch.cuda.synchronize()
    scans, waits = [], []
    for r in range(world):                                   # one process drives both ranks: scan + publish and wait are
        g = torch.cuda.CUDAGraph()                           # captured separately (one rank per process captures one graph)
        with torch.cuda.graph(g, stream=streams[r]):
            contains_bits_sharded(evs[r], wins[r], *local[r], total=totals[r], defer_wait=True)
        scans.append(g)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[r]):
            wait_sharded(wins[r], totals[r])
        waits.append(g)
    torch.cuda.synchronize()
    for rep in range(3):
        for graphs in (scans, waits):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    graphs[r].replay()
        torch.cuda.synchronize()
        for r in range(world):
            wins[r].check()
            assert int(totals[r].item()) == want_count and torch.equal(wins[r].result_bits(), want_bits)
