"""Sharding of the sample axis over the GPUs of one box, and the result gathers.

Every sample / initial state / Monte-Carlo run is independent and every matrix is a tiny constant replicated on each
GPU, so the hot path needs no data-path collective: rank r evaluates the contiguous index range ``shard_range(n, r,
world)`` (whole bitset words per rank).  ``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests) is
used only to collect results: the membership bitset words, member counts, and the per-sample QP outputs.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

from .grids import shard_range


def _dist():
    import torch.distributed as dist
    return dist


def world_info() -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def padded_shard_len(n: int, world: int, align: int = 32) -> int:
    """Samples per rank before clipping to n (the same on every rank, a multiple of ``align``)."""
    per = -(-n // world)
    return -(-per // align) * align


def gather_bitset(local_bits, n: int, async_op: bool = False, out=None, dst: Optional[int] = None):
    """Gather the per-rank bitset words of an n-sample set sharded with ``shard_range(n, rank, world)``.

    ``local_bits``: int32 tensor with the words of this rank's range (ceil(len / 32) words).  With ``dst=None`` every
    rank receives the full bitset (all-gather); with ``dst=r`` only rank r does (gather: 1/world of the traffic per
    sender).  Returns the full bitset (ceil(n / 32) int32 words; ``None`` on the ranks that do not receive) - or
    ``(handle, finish)`` when ``async_op`` - where ``finish()`` trims the padding."""
    import torch
    dist = _dist()
    rank, world = world_info()
    words_total = (n + 31) // 32
    if world == 1:
        res = local_bits[:words_total]
        return (None, lambda: res) if async_op else res
    per_words = padded_shard_len(n, world) // 32
    send = local_bits
    if send.numel() != per_words:                       # last ranks own fewer (or zero) words: pad with zeros
        send = torch.zeros(per_words, dtype=local_bits.dtype, device=local_bits.device)
        send[:local_bits.numel()] = local_bits
    if dst is None:
        if out is None:
            out = torch.empty(world * per_words, dtype=local_bits.dtype, device=local_bits.device)
        handle = dist.all_gather_into_tensor(out, send, async_op=async_op)
    else:
        if rank == dst and out is None:
            out = torch.empty(world * per_words, dtype=local_bits.dtype, device=local_bits.device)
        parts = list(out.view(world, per_words).unbind(0)) if rank == dst else None
        handle = dist.gather(send, parts, dst=dst, async_op=async_op)

    def finish():
        return out[:words_total] if out is not None and (dst is None or rank == dst) else None

    return (handle, finish) if async_op else finish()


def reduce_count(local_count):
    """Sum of the per-rank member counts (int64 tensor of one element), on every rank."""
    rank, world = world_info()
    if world > 1:
        _dist().all_reduce(local_count)
    return local_count


def gather_samples(local: Dict[str, "object"], n: int, sample_dim: Dict[str, int]) -> Dict[str, "object"]:
    """All-gather per-sample result tensors of an n-sample batch sharded with ``shard_range``.

    ``local[k]`` holds this rank's samples along dimension ``sample_dim[k]`` (e.g. ``u0`` (2, B_r) -> dim 1,
    ``status`` (B_r,) -> dim 0).  Returns tensors with all n samples on every rank."""
    import torch
    dist = _dist()
    rank, world = world_info()
    if world == 1:
        return dict(local)
    per = padded_shard_len(n, world)
    out = {}
    for key, t in local.items():
        d = sample_dim[key]
        t = t.movedim(d, 0).contiguous()
        pad_shape = (per,) + tuple(t.shape[1:])
        send = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        send[:t.shape[0]] = t
        recv = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(recv, send)
        # rank r's valid samples are recv[r * per : r * per + len_r]; ranges are contiguous, so the first n rows of
        # the concatenation are exactly samples 0..n-1 when per divides the offsets (it does: lo_r = r * per)
        out[key] = recv[:n].movedim(0, d).contiguous()
    return out


class ShardedTerminalSet:
    """Membership of a sample set that is partitioned over the ranks: each rank scans its own range with the CUDA
    kernel; ``contains_bits`` returns the local words, ``gather`` the full bitset and the global count."""

    def __init__(self, evaluator):
        self.evaluator = evaluator

    def local_range(self, n: int) -> Tuple[int, int]:
        rank, world = world_info()
        return shard_range(n, rank, world)

    def contains_bits_local(self, x, y, psi, v, mode: int = 1):
        return self.evaluator.contains_bits(x, y, psi, v, mode=mode)

    def gather(self, local_bits, local_count, n: int):
        return gather_bitset(local_bits, n), reduce_count(local_count.clone())
