"""Multi-GPU plumbing on the CPU: the shard partition, and the result gathers over a world_size-2 (and 3) gloo group.
The local results are produced by the oracle here (test infrastructure); on the box they come from the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 10 ** 8, 10 ** 6 + 17])
@pytest.mark.parametrize("world", [1, 2, 4, 8, 3])
def test_shard_ranges_partition_the_samples_on_word_boundaries(n, world):
    from carmpc_b200.grids import shard_range
    from carmpc_b200.sharding import padded_shard_len
    prev = 0
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        assert lo == prev and lo <= hi <= n
        assert lo % 32 == 0 or lo == n
        assert hi - lo <= padded_shard_len(n, world)
        assert lo == min(r * padded_shard_len(n, world), n)
        prev = hi
    assert prev == n


def _worker(rank, world, port, n, tmp):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from carmpc_b200.grids import shard_range
        from carmpc_b200.sharding import gather_bitset, reduce_count, gather_samples
        from oracle import c_oracle
        Ab = np.load(os.path.join(GOLDEN, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy"))
        rng = np.random.default_rng(0)                       # same samples on every rank
        p = np.array([30, 1.5, 0, 0]) + rng.uniform(-1, 1, size=(n, 4)) * np.array([10.0, 2.0, 0.4, 3.0])
        lo, hi = shard_range(n, rank, world)
        bits, cnt = c_oracle.membership_bits(Ab, *[np.ascontiguousarray(c) for c in p[lo:hi].T], threads=1)
        full = gather_bitset(torch.from_numpy(bits.view(np.int32)), n)
        total = reduce_count(torch.tensor([cnt], dtype=torch.int64))
        want_bits, want_cnt = c_oracle.membership_bits(Ab, *[np.ascontiguousarray(c) for c in p.T], threads=1)
        assert np.array_equal(full.numpy().view(np.uint32), want_bits)
        assert int(total.item()) == want_cnt
        # gather to one rank only
        only0 = gather_bitset(torch.from_numpy(bits.view(np.int32)), n, dst=0)
        if rank == 0:
            assert np.array_equal(only0.numpy().view(np.uint32), want_bits)
        else:
            assert only0 is None
        # async form
        handle, finish = gather_bitset(torch.from_numpy(bits.view(np.int32)), n, async_op=True)
        if handle is not None:
            handle.wait()
        assert np.array_equal(finish().numpy().view(np.uint32), want_bits)
        # per-sample results (QP layout: u0 (2, B), status (B,))
        local = {"u0": torch.from_numpy(np.ascontiguousarray(p[lo:hi, :2].T)),
                 "status": torch.from_numpy((p[lo:hi, 3] > 0).astype(np.int32))}
        got = gather_samples(local, n, {"u0": 1, "status": 0})
        assert np.array_equal(got["u0"].numpy(), p[:, :2].T)
        assert np.array_equal(got["status"].numpy(), (p[:, 3] > 0).astype(np.int32))
        open(os.path.join(tmp, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 100_003), (2, 40), (3, 1_000)])
def test_gathers_over_gloo(world, n, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_group_aligned_shards_tile_the_sample_set():
    """The fused (peer-window) scans shard on whole 1024-sample groups: one 128-byte bitset line per group."""
    from carmpc_b200.grids import shard_range
    for n in (0, 1, 1023, 1024, 1025, 10 ** 8, 3_000_077):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                lo, hi = shard_range(n, r, world, align=1024)
                assert lo == covered and lo % 1024 == 0 or lo == n
                assert hi >= lo
                covered = hi
            assert covered == n
    lens = [shard_range(10 ** 8, r, 8, align=1024) for r in range(8)]
    assert max(h - l for l, h in lens) - min(h - l for l, h in lens) <= 8 * 1024


def test_cyclic_layout_partitions_the_sample_set():
    """PeerWindow's cyclic layout (group g of 1024 samples belongs to rank g % world): the ranks' index sets partition
    [0, n), counts match, only the owner of the last group holds a partial one.  Host logic only (no window is created)."""
    import torch
    from carmpc_b200.sharding import PeerWindow
    for n in (0, 1, 1023, 1024, 1025, 5000, 3_000_077, 10 ** 6):
        for world in (1, 2, 3, 4, 8):
            parts = []
            for r in range(world):
                w = PeerWindow.__new__(PeerWindow)
                w.n_total, w.rank, w.world, w.layout = n, r, world, "cyclic"
                idx = w.local_index()
                assert idx.numel() == w.local_count()
                assert w._placement() == (r * 1024, world)
                parts.append(idx)
            allidx = torch.cat(parts)
            assert allidx.numel() == n
            if n:
                assert torch.equal(torch.sort(allidx).values, torch.arange(n))
            counts = [p.numel() for p in parts]
            assert max(counts) - min(counts) <= 1024
