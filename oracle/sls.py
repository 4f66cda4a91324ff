"""NumPy restatement of the fast-SLS numerical kernels of the reference.

Follows (reference file:line):
  * ``solver/fast_SLS_jit.py:43-62``   riccati_step_njit
  * ``solver/fast_SLS_jit.py:65-84``   _backward_solve_numba
  * ``solver/fast_SLS_jit.py:87-117``  _propagate
  * ``solver/fast_SLS_jit.py:120-188`` _backoff_from_phi
  * ``solver/fast_SLS_jit.py:475-487`` evaluate_dual_eta
Operation order inside each step is kept identical to the reference (same products, same association) so that
differences against the reference's Numba build are at the level of BLAS summation order only.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np


def riccati_step(A, B, Cx, Cu, Sk):
    """solver/fast_SLS_jit.py:43-62."""
    x = B.T @ Sk
    y = A.T @ Sk
    H = Cu + x @ B
    F = x @ A
    K = -np.linalg.solve(H, F)
    S = Cx + y @ (A + B @ K)
    S = 0.5 * (S + S.T)
    return K, S


def backward_solve(N, nx, nu, A, B, G, Gf, eta, eta_f, Q_reg, R_reg, Q_reg_f):
    """solver/fast_SLS_jit.py:65-84.  eta: (N, N, ni); eta_f: (N+1, ni_f).  Column j = N never reads eta (Q8)."""
    S = np.zeros((N + 1, N + 1, nx, nx))
    K = np.zeros((N, N + 1, nu, nx))
    for jj in range(N + 1):
        S[N, jj] = Gf.T @ (eta_f[jj].reshape(-1, 1) * Gf) + Q_reg_f
        for kk in range(N - 1, jj - 1, -1):
            C = G.T @ (eta[kk, jj].reshape(-1, 1) * G)
            Cx = C[:nx, :nx] + Q_reg
            Cu = C[nx:, nx:] + R_reg
            K[kk, jj], S[kk, jj] = riccati_step(A[kk], B[kk], Cx, Cu, S[kk + 1, jj])
    return S, K


def propagate(A, B, E, K):
    """solver/fast_SLS_jit.py:87-117."""
    N, nx, nu, nw = A.shape[0], A.shape[1], B.shape[2], E.shape[2]
    Phi_x = np.zeros((N + 1, N + 1, nx, nw))
    Phi_u = np.zeros((N, N + 1, nu, nw))
    for j in range(N + 1):
        Phi_x[j, j] = E[j]
    for kk in range(N):
        for jj in range(kk + 1):
            Phi_u[kk, jj] = K[kk, jj] @ Phi_x[kk, jj]
            Phi_x[kk + 1, jj] = (A[kk] + B[kk] @ K[kk, jj]) @ Phi_x[kk, jj]
    return Phi_x, Phi_u


def backoff_from_phi(Phi_x, Phi_u, Gx, Gu, Gf, epsilon):
    """solver/fast_SLS_jit.py:120-188."""
    N, ni, ni_f = Phi_u.shape[0], Gx.shape[0], Gf.shape[0]
    beta = np.zeros((N, N, ni))
    beta_f = np.zeros((N + 1, ni_f))
    for kk in range(N):
        for jj in range(kk + 1):
            Z = Gx @ Phi_x[kk, jj] + Gu @ Phi_u[kk, jj]
            beta[kk, jj] = np.maximum((Z * Z).sum(axis=1), epsilon)
    for jj in range(N + 1):
        Zf = Gf @ Phi_x[N, jj]
        beta_f[jj] = np.maximum((Zf * Zf).sum(axis=1), epsilon)
    backoff = np.zeros((N, ni))
    for kk in range(N):
        backoff[kk] = np.sqrt(beta[kk, : kk + 1]).sum(axis=0)
    backoff_f = np.sqrt(beta_f).sum(axis=0)
    return beta, beta_f, backoff, backoff_f


def evaluate_dual_eta(N, mu, mu_f, beta, beta_f, eps, eta=None, eta_f=None):
    """solver/fast_SLS_jit.py:475-487.  mu: (ni, N); entries eta[k, j] with j > k are left untouched (zero)."""
    ni, ni_f = mu.shape[0], mu_f.shape[0]
    b = np.maximum(beta, eps)
    bf = np.maximum(beta_f, eps)
    eta = np.zeros((N, N, ni)) if eta is None else eta
    eta_f = np.zeros((N + 1, ni_f)) if eta_f is None else eta_f
    for jj in range(N):
        for kk in range(jj, N):
            eta[kk, jj] = mu[:, kk] / (2.0 * np.sqrt(b[kk, jj]))
    for jj in range(N + 1):
        eta_f[jj] = mu_f / (2.0 * np.sqrt(bf[jj]))
    return eta, eta_f
