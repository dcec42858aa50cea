"""Synthetic sample grids of the BASELINE configurations, as structure-of-arrays float64.

The reference builds its sample grid with ``np.linspace`` / ``np.meshgrid`` (``lib/terminal_set.py:96-106``);
the axes here are produced by the same numpy calls on the host (a few hundred numbers), so every coordinate is
bit-identical to what the reference would compute, and only the tensor-product expansion happens on the device.
"""
from __future__ import annotations

import numpy as np


def config2_axes(points_per_axis: int = 100):
    """RoadMultipleCarsEnv box, slightly larger than the constraint box so every face is crossed; C order
    (x slowest ... v fastest)."""
    k = points_per_axis
    return [np.linspace(5.0, 55.0, k), np.linspace(-3.2, 3.2, k), np.linspace(-0.42, 0.42, k),
            np.linspace(-1.2, 5.2, k)]


def config3_axes(nx: int = 100, ny: int = 100, npsi: int = 10, nv: int = 10):
    """RoadOneCarEnv region-of-attraction grid: 100 x 100 x 10 x 10 = 10^6 initial states."""
    return [np.linspace(5.0, 30.0, nx), np.linspace(-3.0, 3.0, ny), np.linspace(-np.pi / 8, np.pi / 8, npsi),
            np.linspace(-1.0, 5.0, nv)]


def grid_size(axes) -> int:
    return int(np.prod([len(a) for a in axes]))


def materialise_grid(axes, device="cuda", start: int = 0, stop: int | None = None):
    """Expand the tensor grid (C order) into four float64 SoA tensors for flat indices [start, stop)."""
    import torch
    dims = [len(a) for a in axes]
    n = grid_size(axes)
    stop = n if stop is None else stop
    idx = torch.arange(start, stop, device=device, dtype=torch.int64)
    out = []
    stride = n
    for a, d in zip(axes, dims):
        stride //= d
        ax = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
        out.append(ax[(idx // stride) % d].contiguous())
    return out


def materialise_grid_at(axes, index):
    """Four float64 SoA tensors with the grid points (C order) at the flat indices ``index`` (int64 tensor, any device)."""
    import torch
    dims = [len(a) for a in axes]
    stride = grid_size(axes)
    out = []
    for a, d in zip(axes, dims):
        stride //= d
        ax = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(index.device)
        out.append(ax[torch.div(index, stride, rounding_mode="floor") % d].contiguous())
    return out


def materialise_grid_host(axes, start: int = 0, stop: int | None = None):
    """numpy version of ``materialise_grid`` (host SoA arrays, used for the end-to-end host path)."""
    dims = [len(a) for a in axes]
    n = grid_size(axes)
    stop = n if stop is None else stop
    idx = np.arange(start, stop, dtype=np.int64)
    out = []
    stride = n
    for a, d in zip(axes, dims):
        stride //= d
        out.append(np.ascontiguousarray(np.asarray(a, dtype=np.float64)[(idx // stride) % d]))
    return out


def shard_range(n: int, rank: int, world: int, align: int = 32):
    """Contiguous index range of ``rank``: ceil(n / world) rounded up to a multiple of ``align`` samples, so every
    rank owns whole bitset words."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def lattice_seeds(dims, block=(2, 8, 1, 1), start: int = 0, stop: int | None = None) -> np.ndarray:
    """Seed map for ``BatchQP.solve(seed=...)`` on a C-order tensor grid: the grid is cut into blocks of ``block`` points
    per axis and the centre point of each block is its anchor (solved cold); every other point of the block starts
    from the anchor's certified active set.  Returns int32 indices local to the slice [start, stop); a point whose
    anchor falls outside the slice is its own anchor."""
    dims = [int(d) for d in dims]
    n = int(np.prod(dims))
    stop = n if stop is None else stop
    idx = np.arange(start, stop, dtype=np.int64)
    anchor = np.zeros_like(idx)
    stride = n
    for d, b in zip(dims, block):
        stride //= d
        k = (idx // stride) % d
        b = max(1, int(b))
        centre = np.minimum((k // b) * b + b // 2, d - 1)
        anchor += centre * stride
    inside = (anchor >= start) & (anchor < stop)
    local = np.where(inside, anchor - start, idx - start)
    return local.astype(np.int32)
