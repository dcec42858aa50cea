"""Batched entry points over the CUDA library: the sampled forms of the reference's two hot loops.

* ``TerminalSetEvaluator`` - membership of very many states in an H-rep set ``{x : A x <= b}``
  (the test ``lib/terminal_set.py:107-113`` applies to one grid point at a time).
* ``RolloutEvaluator`` - the same set in its sampled LQR-rollout form (``lib/terminal_set.py:53-59,
  198-200``): constraint rows checked along x(t+1) = A_k x(t).
* ``BatchQP`` - one condensed MPC QP per initial state (``lib/mpc.py:318-335`` / ``:461-478``) and
  Monte-Carlo closed loops against the nonlinear plant (``lib/simulator.py:51-69``).

Device entry points take and return torch CUDA tensors (torch is only the allocator / stream owner);
``*_host`` entry points take numpy arrays and go through the library's own overlapped copy pipeline.
Nothing here computes on the CPU: without the CUDA library every call raises ``CarmpcError``.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _capi
from ._capi import CarmpcError, check


def _torch():
    import torch
    return torch


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _stream_ptr(stream=None):
    torch = _torch()
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def unpack_bits(bits: np.ndarray, n: int) -> np.ndarray:
    """uint32 bitset -> bool[n]: bit (i & 31) of word (i >> 5) is sample i."""
    return np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), bitorder="little")[:n].astype(bool)


class _Handle:
    def __init__(self):
        self._h = ctypes.c_void_p()
        self._lib = _capi.load()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.carmpc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_soa(x, y, psi, v):
    torch = _torch()
    n = x.numel()
    for t in (x, y, psi, v):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == n):
            raise ValueError("x, y, psi, v must be contiguous float64 CUDA tensors of equal length")
    return n


class TerminalSetEvaluator(_Handle):
    """``A x <= b`` for every sample; results as a bitset (1 bit per sample, warp-ballot packed)."""

    def __init__(self, A: np.ndarray, b: Optional[np.ndarray] = None):
        super().__init__()
        A = _f64(A)
        Ab = A if b is None else np.hstack((A, _f64(b).reshape(-1, 1)))
        if Ab.ndim != 2 or Ab.shape[1] != 5:
            raise ValueError("expected A (rows, 4) and b (rows,), or [A | b] (rows, 5)")
        self.Ab = np.ascontiguousarray(Ab)
        self.rows = len(Ab)
        check(self._lib.carmpc_polytope_create(_capi.ptr(self.Ab), self.rows, ctypes.byref(self._h)))

    @classmethod
    def from_file(cls, path: str) -> "TerminalSetEvaluator":
        """Load a ``terminal_sets/<env>_<goal>.npy`` H-rep (``[A | b]`` rows)."""
        return cls(np.load(path))

    # ---- device tensors in, device bitset out ------------------------------------------------
    def contains_bits(self, x, y, psi, v, mode: int = 1, bits=None, count=None, stream=None):
        """Returns ``(bits int32[ceil(n/32)], count int64[1])`` CUDA tensors; nothing is synchronised."""
        torch = _torch()
        n = _check_soa(x, y, psi, v)
        if bits is None:
            bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=x.device)
        if count is None:
            count = torch.empty(1, dtype=torch.int64, device=x.device)
        check(self._lib.carmpc_membership_bitset(self._h, x.data_ptr(), y.data_ptr(), psi.data_ptr(), v.data_ptr(),
                                                 n, bits.data_ptr(), count.data_ptr(), mode, _stream_ptr(stream)))
        return bits, count

    def contains_grid_bits(self, axes, axis_to_state=(0, 1, 2, 3), bits=None, count=None, stream=None):
        """Membership on the implicit tensor grid ``axes[0] x axes[1] x axes[2] x axes[3]`` (C order); axis k carries
        state component ``axis_to_state[k]``.  No coordinate array is materialised."""
        torch = _torch()
        axes = [_f64(a).ravel() for a in axes]
        dims = (ctypes.c_int32 * 4)(*[len(a) for a in axes])
        a2s = (ctypes.c_int32 * 4)(*axis_to_state)
        n = int(np.prod([len(a) for a in axes]))
        cat = np.ascontiguousarray(np.concatenate(axes))
        if bits is None:
            bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
        if count is None:
            count = torch.empty(1, dtype=torch.int64, device="cuda")
        check(self._lib.carmpc_membership_grid(self._h, _capi.ptr(cat), dims, a2s, bits.data_ptr(), count.data_ptr(),
                                               _stream_ptr(stream)))
        torch.cuda.current_stream().synchronize()      # `cat` is pageable host memory read by an async copy
        return bits, count

    # ---- host arrays in, host results out ---------------------------------------------------------
    def contains_bits_host(self, x, y, psi, v, mode: int = 1):
        """numpy SoA in -> ``(bits uint32[ceil(n/32)], count)``; copies overlap the kernel inside the library."""
        x, y, psi, v = _f64(x), _f64(y), _f64(psi), _f64(v)
        n = len(x)
        if not (len(y) == len(psi) == len(v) == n):
            raise ValueError("x, y, psi, v must have equal length")
        bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        cnt = ctypes.c_int64(0)
        check(self._lib.carmpc_membership_bitset_host(self._h, _capi.ptr(x), _capi.ptr(y), _capi.ptr(psi),
                                                      _capi.ptr(v), n, _capi.ptr(bits), ctypes.byref(cnt), mode))
        return bits, int(cnt.value)

    def contains_host(self, x, y, psi, v, mode: int = 1) -> np.ndarray:
        """bool[n] membership of numpy SoA samples."""
        bits, _ = self.contains_bits_host(x, y, psi, v, mode)
        return unpack_bits(bits, len(np.atleast_1d(x)))


class RolloutEvaluator(_Handle):
    """Sampled form of the terminal set: e(0) = p - goal, e(t+1) = A_k e(t); state rows for t = 0..k_steps, input
    rows at t = 0 only (``input_every_step=False``, what the reference's construction does) or at every step."""

    def __init__(self, A_k, A_con, b_con, A_in, b_in, goal, k_steps: int, input_every_step: bool = False):
        super().__init__()
        self.A_k, self.A_con, self.b_con = _f64(A_k), _f64(A_con).reshape(-1, 4), _f64(b_con).ravel()
        self.A_in, self.b_in, self.goal = _f64(A_in).reshape(-1, 4), _f64(b_in).ravel(), _f64(goal).ravel()
        self.k_steps = int(k_steps)
        check(self._lib.carmpc_rollout_create(_capi.ptr(self.A_k), _capi.ptr(self.A_con), _capi.ptr(self.b_con),
                                              len(self.b_con), _capi.ptr(self.A_in), _capi.ptr(self.b_in),
                                              len(self.b_in), _capi.ptr(self.goal), self.k_steps,
                                              1 if input_every_step else 0, ctypes.byref(self._h)))

    @classmethod
    def from_env(cls, env, k_steps: int, input_every_step: bool = False) -> "RolloutEvaluator":
        """Rows exactly as ``calc_terminal_set`` builds them: unit-norm state rows shifted to the goal, and
        ``[I; -I] K`` input rows (``lib/terminal_set.py:145-161, 198-200``)."""
        from .lib.terminal_set import lqr_closed_loop
        from .lib import polytope_ops as pc
        _, A_k, A_con, b_con, A_in, b_in = lqr_closed_loop(env)
        goal = np.array(env.goal, dtype=float)
        pc_state = pc.Polytope(A_con, b_con).translation(-goal)
        pc_in = pc.Polytope(A_in, b_in)
        return cls(A_k, pc_state.A, pc_state.b, pc_in.A, pc_in.b, goal, k_steps, input_every_step)

    def contains_bits(self, x, y, psi, v, want_first_violation: bool = False, bits=None, count=None, stream=None):
        torch = _torch()
        n = _check_soa(x, y, psi, v)
        if bits is None:
            bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=x.device)
        if count is None:
            count = torch.empty(1, dtype=torch.int64, device=x.device)
        first = torch.empty(n, dtype=torch.int32, device=x.device) if want_first_violation else None
        check(self._lib.carmpc_rollout_bitset(self._h, x.data_ptr(), y.data_ptr(), psi.data_ptr(), v.data_ptr(), n,
                                              bits.data_ptr(), first.data_ptr() if first is not None else None,
                                              count.data_ptr(), _stream_ptr(stream)))
        return (bits, count, first) if want_first_violation else (bits, count)

    def contains_bits_host(self, x, y, psi, v, want_first_violation: bool = False):
        x, y, psi, v = _f64(x), _f64(y), _f64(psi), _f64(v)
        n = len(x)
        bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        first = np.zeros(n, dtype=np.int32) if want_first_violation else None
        cnt = ctypes.c_int64(0)
        check(self._lib.carmpc_rollout_bitset_host(self._h, _capi.ptr(x), _capi.ptr(y), _capi.ptr(psi), _capi.ptr(v), n,
                                                   _capi.ptr(bits), _capi.ptr(first), ctypes.byref(cnt)))
        return (bits, int(cnt.value), first) if want_first_violation else (bits, int(cnt.value))


def measure_peak(which: str) -> float:
    """Device micro-benchmarks: 'fp32' / 'fp64' FMA TFLOP/s, 'hbm' copy GB/s (read + write)."""
    idx = {"fp32": 0, "fp64": 1, "hbm": 2}[which]
    val = ctypes.c_double(0.0)
    check(_capi.load().carmpc_measure_peak(idx, ctypes.byref(val)))
    return float(val.value)
