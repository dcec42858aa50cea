"""Fourier-Motzkin projection of the admissible set {(x0, u) : G x0 + H u + phi <= 0} onto x0.

Same two entry points as the reference's ``lib/in_adm_set.py`` (``algorithm_1`` :4-40 eliminates one
input, ``algorithm_2`` :43-77 eliminates all of them).  Nothing in the reference calls them; their
sampled equivalent is the per-state QP feasibility flag computed on the GPU.  Row order of the
result is the reference's: rows with a zero coefficient first, then every (positive, negative) pair.
"""
import numpy as np


def algorithm_1(G: np.ndarray, H: np.ndarray, phi: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Eliminate a single input.  G (s, n), H (s,), phi (s,) -> P (r, n), gamma (r,)."""
    G = np.asarray(G, dtype=float)
    H = np.asarray(H, dtype=float)
    phi = np.asarray(phi, dtype=float)
    zero = np.flatnonzero(H == 0)
    pos = np.flatnonzero(H > 0)
    neg = np.flatnonzero(H < 0)
    assert len(G) == len(zero) + len(pos) + len(neg)

    C = np.column_stack((G, phi))
    ii, jj = np.meshgrid(pos, neg, indexing='ij')
    ii, jj = ii.ravel(), jj.ravel()
    pairs = H[ii, None] * C[jj] - H[jj, None] * C[ii]
    D = np.vstack((C[zero], pairs))
    return D[:, :-1], D[:, -1]


def algorithm_2(G: np.ndarray, H: np.ndarray, phi: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Eliminate all m > 1 inputs, last column first.  Returns P x0 + gamma <= 0."""
    G = np.asarray(G, dtype=float)
    H = np.asarray(H, dtype=float)
    m = H.shape[1] if H.ndim >= 2 else 1
    assert m > 1, "Use algorithm_1"
    aug = np.column_stack((G, H[:, :m - 1]))
    last = H[:, m - 1]
    gamma = np.asarray(phi, dtype=float)
    for _ in range(m):
        P, gamma = algorithm_1(aug, last, gamma)
        aug, last = P[:, :-1], P[:, -1]
    return P, gamma
