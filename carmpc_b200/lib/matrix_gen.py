"""Prediction and cost matrices of the condensed MPC problem.

Same functions, argument order and return values as the reference's ``lib/matrix_gen.py``
(``predmod`` :6-32, ``costgen`` :35-72, ``stack_matrix_along_diag`` :75-82).  Host-side, executed
once per controller; the results are staged to the GPU by ``carmpc_b200.condensed``.

The reference calls ``np.linalg.matrix_power`` O(N^2) times; here the powers A^k and the impulse
responses A^k B are built by one recurrence each and scattered into the block-Toeplitz S.
"""
import numpy as np


def predmod(A: np.ndarray, B: np.ndarray, N: int) -> tuple[np.ndarray, np.ndarray]:
    """x_ = T x0 + S u_ for x(k+1) = A x(k) + B u(k), x_ = {x(0)..x(N)}, u_ = {u(0)..u(N-1)}.

    T = [I; A; ...; A^N]                       ((N+1)*nx, nx)
    S = block lower-triangular Toeplitz, block (i, j) = A^(i-j-1) B for i > j, first block row zero.
    """
    A = np.asarray(A, dtype=float)
    B = np.asarray(B, dtype=float)
    nx, nu = B.shape
    powers = np.empty((N + 1, nx, nx))
    powers[0] = np.eye(nx)
    for k in range(1, N + 1):
        # agreement with the reference's matrix_power products is checked to 1e-12 in
        # tests/test_host_matrices.py
        powers[k] = powers[k - 1] @ A
    T = powers.reshape((N + 1) * nx, nx).copy()

    impulse = powers[:N] @ B                     # impulse[k] = A^k B
    S = np.zeros(((N + 1) * nx, N * nu))
    for j in range(N):
        S[(j + 1) * nx:, j * nu:(j + 1) * nu] = impulse[:N - j].reshape(-1, nu)
    return T, S


def stack_matrix_along_diag(A: np.ndarray, N: int) -> np.ndarray:
    """blkdiag(A, ..., A) with N copies."""
    return np.kron(np.eye(N), np.asarray(A, dtype=float))


def costgen(Q: np.ndarray, R: np.ndarray, P: np.ndarray, T: np.ndarray, S: np.ndarray,
            nx: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """V_N = u_' H u_ + 2 (h x0)' u_ + x0' const x0 with terminal weight P on x(N).

    H = R_hat + S' Q_hat S,  h = S' Q_hat T,  const = T' Q_hat T  (reference :66-70).
    """
    N = len(S) // nx - 1
    Q_hat = stack_matrix_along_diag(Q, N + 1)
    Q_hat[N * nx:, N * nx:] = P
    R_hat = stack_matrix_along_diag(R, N)
    QT = Q_hat @ T
    H = R_hat + S.T @ (Q_hat @ S)
    h = S.T @ QT
    constant = T.T @ QT
    return H, h, constant
