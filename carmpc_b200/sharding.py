"""Sharding of the sample axis over the GPUs of one box, and the result gathers.

Every sample / initial state / Monte-Carlo run is independent and every matrix is a tiny constant replicated on each
GPU, so the hot path needs no data-path collective: rank r evaluates the contiguous index range ``shard_range(n, r,
world)`` (whole bitset words per rank).  ``torch.distributed`` (NCCL over NVLink on the box, gloo in the CPU tests) is
used only to collect results: the membership bitset words, member counts, and the per-sample QP outputs.

Two ways to collect the membership bitsets:
  * ``gather_bitset`` / ``reduce_count`` - NCCL all-gather / all-reduce after the scan (any backend; gloo in the CPU tests);
  * ``PeerWindow`` + ``contains_bits_sharded`` - the scan kernel itself writes every rank's copy of the bitset through
    NVLink peer mappings (CUDA IPC), and a one-warp kernel exchanges counts and completion flags: the all-gather is
    fused into the kernel, one collective step is two launches and no NCCL call (what ``bench.py --gpus N`` times).
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


def gather_bitset(local_bits, n: int, async_op: bool = False, out=None, dst: Optional[int] = None, shard_align: int = 32):
    """Gather the per-rank bitset words of an n-sample set sharded with ``shard_range(n, rank, world)``.

    ``local_bits``: int32 tensor with the words of this rank's range (ceil(len / 32) words).  With ``dst=None`` every
    rank receives the full bitset (all-gather); with ``dst=r`` only rank r does (gather: 1/world of the traffic per
    sender).  Returns the full bitset (ceil(n / 32) int32 words; ``None`` on the ranks that do not receive) - or
    ``(handle, finish)`` when ``async_op`` - where ``finish()`` trims the padding.  ``shard_align``: the alignment the
    set was sharded with (``shard_range(..., align=shard_align)``; 1024 for the group-aligned shards of ``PeerWindow``)."""
    import torch
    dist = _dist()
    rank, world = world_info()
    words_total = (n + 31) // 32
    if world == 1:
        res = local_bits[:words_total]
        return (None, lambda: res) if async_op else res
    per_words = padded_shard_len(n, world, shard_align) // 32
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


class PeerWindow:
    """This rank's window of a sample set sharded over the GPUs of one box (``carmpc_shard_*``): the full bitset
    (double-buffered), per-rank member counts and per-rank step flags in peer-mapped device memory.  The scan kernels
    write their words into every rank's window over NVLink; no collective call gathers them afterwards.

    ``PeerWindow(n)`` under ``torch.distributed`` (one process per GPU) exchanges the CUDA IPC handles with
    ``all_gather_object``; ``PeerWindow.local_group(n, world)`` builds all windows of a group inside one process
    (several ranks on one or more devices: tests, single-process multi-GPU drivers)."""

    GROUP = 1024        # samples per bitset line (32 words): the unit the scan kernel hands to a warp

    def __init__(self, n_total: int, rank: Optional[int] = None, world: Optional[int] = None, _connect: bool = True,
                 layout: str = "cyclic"):
        import ctypes
        if layout not in ("cyclic", "contiguous"):
            raise ValueError("layout must be 'cyclic' or 'contiguous'")
        self.layout = layout
        from . import _capi
        self._lib = _capi.load()
        r, w = world_info()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.n_total = int(n_total)
        self._h = ctypes.c_void_p()
        _capi.check(self._lib.carmpc_shard_create(self.rank, self.world, self.n_total, ctypes.byref(self._h)))
        if _connect:
            self._connect_ipc()

    def _connect_ipc(self):
        import ctypes
        from . import _capi
        if self.world == 1:
            return
        mine = (ctypes.c_ubyte * 64)()
        _capi.check(self._lib.carmpc_shard_export(self._h, mine))
        handles = [None] * self.world
        _dist().all_gather_object(handles, bytes(mine))
        blob = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
        _capi.check(self._lib.carmpc_shard_connect(self._h, blob))
        _dist().barrier()                      # nobody writes into a window before every rank has mapped it

    @classmethod
    def local_group(cls, n_total: int, world: int, devices=None, layout: str = "cyclic"):
        """All ``world`` windows in this process; ``devices[r]`` is the CUDA device index of rank r (default: current)."""
        import ctypes
        import torch
        from . import _capi
        wins = []
        for r in range(world):
            if devices is not None:
                with torch.cuda.device(devices[r]):
                    wins.append(cls(n_total, rank=r, world=world, _connect=False, layout=layout))
            else:
                wins.append(cls(n_total, rank=r, world=world, _connect=False, layout=layout))
        arr = (ctypes.c_void_p * world)(*[w._h for w in wins])
        for w in wins:
            _capi.check(w._lib.carmpc_shard_connect_local(w._h, arr))
        return wins

    def shard(self) -> Tuple[int, int]:
        """Contiguous layout: [lo, hi) of this rank, whole 1024-sample groups (one 128-byte bitset line per group and
        destination)."""
        if self.layout != "contiguous":
            raise ValueError("shard() is the contiguous layout; use local_count() / local_index() for the cyclic one")
        return shard_range(self.n_total, self.rank, self.world, align=self.GROUP)

    def local_count(self) -> int:
        """Samples this rank evaluates."""
        if self.layout == "contiguous":
            lo, hi = self.shard()
            return hi - lo
        g = self.GROUP
        n_groups = -(-self.n_total // g)                     # the last group may be partial
        mine = max(0, -(-(n_groups - self.rank) // self.world))
        count = mine * g
        if mine and (self.rank + (mine - 1) * self.world) == n_groups - 1 and self.n_total % g:
            count -= g - self.n_total % g
        return count

    def local_index(self, device="cpu"):
        """int64 tensor: for every local sample its index in the whole set.  Cyclic layout: the 1024-sample groups of
        the set are dealt round-robin (group g belongs to rank g % world), so every rank sees the same mix of cheap and
        expensive regions of a structured sample set."""
        import torch
        k = torch.arange(self.local_count(), dtype=torch.int64, device=device)
        if self.layout == "contiguous":
            return k + self.shard()[0]
        g = self.GROUP
        return (torch.div(k, g, rounding_mode="floor") * self.world + self.rank) * g + k % g

    def _placement(self) -> Tuple[int, int]:
        """(first_sample, group_stride) of carmpc_*_bitset_sharded."""
        if self.layout == "contiguous":
            return self.shard()[0], 1
        return self.rank * self.GROUP, self.world

    def result_bits(self):
        """The full bitset of the last completed step as an int32 CUDA tensor view of the window (valid until the
        second-next collective step)."""
        import ctypes
        import torch
        ptr, steps = ctypes.c_void_p(), ctypes.c_int64(0)
        from . import _capi
        _capi.check(self._lib.carmpc_shard_result(self._h, ctypes.byref(ptr), ctypes.byref(steps)))
        words = (self.n_total + 31) // 32
        if words == 0:
            return torch.empty(0, dtype=torch.int32, device="cuda")
        iface = {"shape": (words,), "typestr": "<i4", "data": (ptr.value, False), "version": 2}
        holder = type("_WindowView", (), {"__cuda_array_interface__": iface, "_keep": self})()
        return torch.as_tensor(holder, device="cuda")

    def check(self) -> None:
        from . import _capi
        _capi.check(self._lib.carmpc_shard_check(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.carmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def contains_bits_sharded(evaluator, window: "PeerWindow", x, y, psi, v, mode: int = 1, total=None, stream=None,
                          defer_wait: bool = False):
    """One collective step: membership (``TerminalSetEvaluator``) or rollout form (``RolloutEvaluator``) of this rank's
    shard (``window.local_index()``), bitset words written into every rank's window by the scan kernel itself, member counts
    and completion flags exchanged by a one-warp kernel.  ``total``: int64 CUDA tensor (1,) receiving the global
    member count.  Nothing is synchronised; ``window.result_bits()`` is valid in stream order.

    ``defer_wait``: stop after publishing this rank's flag; ``wait_sharded(window, total)`` then enqueues the wait.  Needed
    when ONE process drives several ranks: enqueue every rank's scan first, then every rank's wait."""
    import ctypes
    import torch
    from . import _capi
    from .batch import RolloutEvaluator, _check_soa
    n = _check_soa(x, y, psi, v) if x.numel() else 0
    if n != window.local_count():
        raise ValueError(f"rank {window.rank} owns {window.local_count()} samples of the set ({window.layout} layout), got {n}")
    lo, stride = window._placement()
    if total is None:
        total = torch.empty(1, dtype=torch.int64, device=x.device)
    s = torch.cuda.current_stream() if stream is None else stream
    st = ctypes.c_void_p(s.cuda_stream)
    if isinstance(evaluator, RolloutEvaluator):
        _capi.check(evaluator._lib.carmpc_rollout_bitset_sharded(evaluator._h, window._h, x.data_ptr(), y.data_ptr(),
                                                                 psi.data_ptr(), v.data_ptr(), n, lo, stride, total.data_ptr(),
                                                                 int(defer_wait), st))
    else:
        _capi.check(evaluator._lib.carmpc_membership_bitset_sharded(evaluator._h, window._h, x.data_ptr(), y.data_ptr(),
                                                                    psi.data_ptr(), v.data_ptr(), n, lo, stride, mode,
                                                                    total.data_ptr(), int(defer_wait), st))
    return total


def wait_sharded(window: "PeerWindow", total, stream=None):
    """Second half of a collective step started with ``defer_wait=True`` (``carmpc_shard_wait``)."""
    import ctypes
    import torch
    from . import _capi
    s = torch.cuda.current_stream() if stream is None else stream
    _capi.check(window._lib.carmpc_shard_wait(window._h, total.data_ptr(), ctypes.c_void_p(s.cuda_stream)))
    return total


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

    def contains_bits_fused(self, window: PeerWindow, x, y, psi, v, mode: int = 1, total=None):
        """Scan + all-gather in one step (no NCCL call): see ``contains_bits_sharded``."""
        return contains_bits_sharded(self.evaluator, window, x, y, psi, v, mode=mode, total=total)
