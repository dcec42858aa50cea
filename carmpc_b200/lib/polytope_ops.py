"""Minimal H-representation polytope algebra used by the terminal-set construction.

The reference leans on the third-party ``polytope`` package for four operations
(``lib/terminal_set.py:41-43, 64-66, 199-203``): construction with row normalisation, translation,
intersection and redundancy removal.  ``polytope`` is not part of the reference tree (and not
installed); its behaviour for these four operations is restated here on top of
``scipy.optimize.linprog`` so that ``calc_terminal_set`` reproduces the shipped ``terminal_sets/*.npy``
row for row (tests/test_lib_golden.py::test_calc_terminal_set_regenerates_shipped_fixture pins five fixtures to < 1e-9).

Conventions: a polytope is {x : A x <= b}; every constructed polytope has unit-norm rows; rows
whose norm is <= 1e-10 are dropped at construction.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import linprog

ABS_TOL = 1e-7


class Polytope:
    def __init__(self, A: np.ndarray, b: np.ndarray, normalize: bool = True):
        A = np.array(A, dtype=float).reshape(-1, np.shape(A)[-1]) if np.size(A) else np.zeros((0, 0))
        b = np.array(b, dtype=float).reshape(-1)
        if normalize and A.size:
            norms = np.sqrt(np.sum(A * A, axis=1))
            keep = norms > 1e-10
            A, b, norms = A[keep], b[keep], norms[keep]
            A = A / norms[:, None]
            b = b / norms
        self.A, self.b = A, b

    @property
    def dim(self) -> int:
        return self.A.shape[1]

    def translation(self, d) -> "Polytope":
        """{x + d : x in self}:  A (x - d) <= b  <=>  A x <= b + A d."""
        d = np.asarray(d, dtype=float)
        return Polytope(self.A.copy(), self.b + self.A @ d, normalize=False)

    def intersect(self, other: "Polytope") -> "Polytope":
        return reduce(Polytope(np.vstack((self.A, other.A)), np.hstack((self.b, other.b))))

    def contains(self, points: np.ndarray) -> np.ndarray:
        """Membership of the rows of ``points`` (n, dim) with the reference's ``<=`` semantics."""
        return np.all(np.asarray(points) @ self.A.T <= self.b, axis=1)

    def __len__(self):
        return len(self.b)


def _maximise(c: np.ndarray, A: np.ndarray, b: np.ndarray):
    """max c'x s.t. A x <= b.  Returns (status, value); status 0 ok, 3 unbounded, 2 infeasible."""
    res = linprog(-c, A_ub=A, b_ub=b, bounds=[(None, None)] * A.shape[1], method="highs")
    return res.status, (-res.fun if res.status == 0 else None)


def reduce(poly: Polytope, abs_tol: float = ABS_TOL) -> Polytope:
    """Remove redundant rows, keeping the survivors in their original order.

    1. of rows describing the same hyperplane direction (cosine > 1 - abs_tol) keep the tightest;
    2. row k is redundant when  max {a_k x : A x <= b, row k relaxed by 1}  <=  b_k + abs_tol.
    """
    idx = reduce_indices(poly, abs_tol)
    return Polytope(poly.A[idx], poly.b[idx])


def reduce_indices(poly: Polytope, abs_tol: float = ABS_TOL) -> np.ndarray:
    """Indices (into ``poly``'s rows, ascending) of the rows ``reduce`` keeps."""
    A, b = poly.A, poly.b
    finite = np.flatnonzero(b != np.inf)
    A, b = A[finite], b[finite]
    m = len(b)
    if m == 0:
        return finite

    inv_norm = 1.0 / np.sqrt(np.sum(A * A, axis=1))
    unit = A * inv_norm[:, None]
    cosine = unit @ unit.T
    drop = set()
    for i in range(m):
        for j in range(i + 1, m):
            if cosine[i, j] > 1 - abs_tol:
                drop.add(i if b[i] * inv_norm[i] > b[j] * inv_norm[j] else j)
    keep = [k for k in range(m) if k not in drop]
    A, b = A[keep], b[keep]

    survivors = []
    for k in range(len(b)):
        relaxed = b.copy()
        relaxed[k] += 1.0
        status, value = _maximise(A[k], A, relaxed)
        if status == 3 or (status == 0 and value - b[k] > abs_tol):
            survivors.append(k)
    return finite[np.asarray(keep, dtype=np.int64)[survivors]] if survivors else finite[:0]
