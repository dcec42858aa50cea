"""ctypes front of oracle/carmpc_oracle.c (test infrastructure; builds the .so on first use)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcarmpc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "carmpc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libcarmpc_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = ctypes.CDLL(build())
        except OSError:
            _lib = ctypes.CDLL(build(force=True))
        dp, i64, i32 = ctypes.POINTER(ctypes.c_double), ctypes.c_int64, ctypes.c_int
        _lib.oracle_membership.restype = i64
        _lib.oracle_membership.argtypes = [dp, i32, dp, dp, dp, dp, i64, ctypes.POINTER(ctypes.c_uint32), i32]
        _lib.oracle_rollout.restype = i64
        _lib.oracle_rollout.argtypes = [dp, dp, dp, i32, dp, dp, i32, dp, i32, i32, dp, dp, dp, dp, i64,
                                        ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int32), i32]
        _lib.oracle_membership_plain.restype = i64
        _lib.oracle_membership_plain.argtypes = [dp, i32, dp, i64, ctypes.POINTER(ctypes.c_uint8)]
        _lib.oracle_max_threads.restype = i32
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def membership_bits(Ab, x, y, psi, v, threads: int = 0):
    """(bits uint32[ceil(n/32)], count) with the fma-chain contract of the CUDA kernel."""
    Ab, x, y, psi, v = _c(Ab), _c(x), _c(y), _c(psi), _c(v)
    n = len(x)
    bits = np.zeros((n + 31) // 32, dtype=np.uint32)
    cnt = lib().oracle_membership(_dp(Ab), len(Ab), _dp(x), _dp(y), _dp(psi), _dp(v), n,
                                  bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), threads)
    return bits, int(cnt)


def rollout_bits(Ak, Acon, bcon, Ain, bin_, goal, k_steps, input_mode, x, y, psi, v, threads: int = 0):
    Ak, Acon, bcon, Ain, bin_, goal = _c(Ak), _c(Acon), _c(bcon), _c(Ain), _c(bin_), _c(goal)
    x, y, psi, v = _c(x), _c(y), _c(psi), _c(v)
    n = len(x)
    bits = np.zeros((n + 31) // 32, dtype=np.uint32)
    first = np.zeros(n, dtype=np.int32)
    cnt = lib().oracle_rollout(_dp(Ak), _dp(Acon), _dp(bcon), len(bcon), _dp(Ain), _dp(bin_), len(bin_), _dp(goal),
                               k_steps, input_mode, _dp(x), _dp(y), _dp(psi), _dp(v), n,
                               bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                               first.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), threads)
    return bits, first, int(cnt)


def membership_plain(Ab, points):
    Ab, points = _c(Ab), _c(points)
    out = np.zeros(len(points), dtype=np.uint8)
    lib().oracle_membership_plain(_dp(Ab), len(Ab), _dp(points), len(points),
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out.astype(bool)


def unpack_bits(bits: np.ndarray, n: int) -> np.ndarray:
    """uint32 bitset -> bool[n] (bit i & 31 of word i >> 5)."""
    return np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)
