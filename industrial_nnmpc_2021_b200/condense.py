"""One-time host-side setup of the condensed regulator QP and of the operators the CUDA
kernels consume.  NumPy/SciPy float64, like the reference's own constructor code
(/root/reference/lib/linearMPC.py:339-395), but built by a backward block recursion instead
of the reference's dense ``tQ`` ((N+1)Nxa square: 12.8 GB at CDU size) and O(N^2)
``matrix_power`` calls (:416-428, :454, :472).

Notation (reference names): ``P = tB'tQtB + tR + tB'tM + tM'tB``, ``tq = (tB'tQ + tM')tA``
(:467-474) with block rows ``tB_k = [A^{k-1}B ... B 0 ...]``.  With ``Phi_d = A^{d-1}B``,

    Lam_{i,j} = sum_{k>i} (A^{k-1-i})' Q_k Phi_{k-j}      (Q_k = Q, Q_N = Pf)
    Lam_{N-1,j} = Pf Phi_{N-j},   Lam_{i-1,j} = Q Phi_{i-j} + A' Lam_{i,j}
    P_{ij} = B' Lam_{i,j} + [i=j] R + [i>j] M' Phi_{i-j}            (i >= j)
    Gam_{i}   = sum_{k>i} (A^{k-1-i})' Q_k A^k,  Gam_{N-1} = Pf A^N, Gam_{i-1} = Q A^i + A' Gam_i
    tq_i   = B' Gam_i + M' A^i
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def dlqr(A, B, Q, R, M=None):
    """LQR gain/cost-to-go for stage cost x'Qx + 2x'Mu + u'Ru (linearMPC.py:22-40)."""
    if M is None:
        M = np.zeros(B.shape)
        At, Qt = A, Q
    else:
        W = scipy.linalg.solve(R, M.T)
        At, Qt = A - B @ W, Q - M @ W
    Pi = scipy.linalg.solve_discrete_are(At, B, Qt, R)
    K = -scipy.linalg.solve(B.T @ Pi @ B + R, B.T @ Pi @ A + M.T)
    return K, Pi


def dlqe(A, C, Q, R):
    """Steady-state Kalman gain (linearMPC.py:42-48)."""
    P = scipy.linalg.solve_discrete_are(A.T, C.T, Q, R)
    L = scipy.linalg.solve(C @ P @ C.T + R, C @ P).T
    return L, P


def c2d(A, B, sample_time):
    """ZOH discretisation through one augmented matrix exponential (linearMPC.py:50-64)."""
    nx, nu = B.shape
    blk = np.zeros((nx + nu, nx + nu))
    blk[:nx, :nx], blk[:nx, nx:] = A, B
    E = scipy.linalg.expm(blk * sample_time)
    return E[:nx, :nx], E[:nx, nx:]


def state_powers(A, N):
    """[A^0, ..., A^N] as an (N+1, nx, nx) array."""
    nx = A.shape[0]
    pw = np.empty((N + 1, nx, nx))
    pw[0] = np.eye(nx)
    for k in range(N):
        pw[k + 1] = A @ pw[k]
    return pw


def condensed_hessian(A, B, Q, R, M, Pf, N):
    """(P, tq) of the dense regulator QP by backward block recursion; O(N^2 nx^2 nu) flops."""
    nx, nu = B.shape
    n = N * nu
    pw = state_powers(A, N)
    # PhiRev = [Phi_N, ..., Phi_1]; block row k of tB (first k blocks) = PhiRev[:, (N-k)nu:]
    PhiRev = np.empty((nx, n))
    for d in range(1, N + 1):
        PhiRev[:, (N - d) * nu:(N - d + 1) * nu] = pw[d - 1] @ B
    P = np.zeros((n, n))
    tq = np.empty((n, nx))
    Lam = Pf @ PhiRev                                   # Lam_{N-1, j}, j = 0..N-1
    Gam = Pf @ pw[N]
    Bt, At, Mt = B.T, A.T, M.T
    for i in range(N - 1, -1, -1):
        w = (i + 1) * nu
        P[i * nu:w, :w] = Bt @ Lam[:, :w]
        P[i * nu:w, i * nu:w] += R
        tq[i * nu:w, :] = Bt @ Gam + Mt @ pw[i]
        if i > 0:
            row_i = PhiRev[:, (N - i) * nu:]            # [Phi_i ... Phi_1]
            P[i * nu:w, :i * nu] += Mt @ row_i
            Lam = Q @ row_i + At @ Lam[:, :i * nu]
            Gam = Q @ pw[i] + At @ Gam
    iu = np.triu_indices(n, 1)
    P[iu] = P.T[iu]
    return P, tq


def reparameterize(A, B, Q, R, M, Krep):
    """u = Krep x + v for an open-loop unstable A (linearMPC.py:366-382): returns (A, Q, M) of the
    re-parameterised problem in the reference's statement order (R, B and Pf do not change)."""
    A2 = A + B @ Krep
    Q2 = Q + Krep.T @ (R @ Krep)
    Q2 = Q2 + M @ Krep + Krep.T @ M.T
    M2 = Krep.T @ R + M
    return A2, Q2, M2


def input_space_operators(A_cl, B, Krep, N, P_v, tq_v):
    """Re-parameterised QP -> the equivalent BOX QP in the original inputs.

    The reference solves for v with the dense inequality G v <= h(x0),
    G = tE (tK tB_N + I), h = te - tE tK tA_N x0 (:476-493), and maps back u = tK (tA_N x0 + tB_N v) + v
    (:507-509).  That map u = T v + S x0, T = I + tK tB_N (unit lower block triangular, hence
    invertible), S = tK tA_N, is a bijection under which G v <= h becomes the plain box
    lb <= u_k <= ub, so the minimiser is the image of the reference's and

        P_u = T^-T P_v T^-1,      tq_u = T^-T tq_v - P_u S,
        V_v(v*) = V_u(u*) + 1/2 x0'(S'P_u S) x0 - (tq_v x0)'T^-1 S x0

    T^-1 is formed by a triangular solve.  (A_cl = A + B Krep is stable, so T, S stay O(1); what an
    unstable plant costs is the conditioning of P_u, which the caller checks.)  Returns
    (P_u, tq_u, T, S)."""
    nx, nu = B.shape
    n = N * nu
    pw = state_powers(A_cl, N)
    T = np.eye(n)
    S = np.empty((n, nx))
    for k in range(N):
        S[k * nu:(k + 1) * nu] = Krep @ pw[k]
        for j in range(k):
            T[k * nu:(k + 1) * nu, j * nu:(j + 1) * nu] = Krep @ (pw[k - j - 1] @ B)
    Ti = scipy.linalg.solve_triangular(T, np.eye(n), lower=True, unit_diagonal=True, check_finite=False)
    P_u = Ti.T @ P_v @ Ti
    P_u = 0.5 * (P_u + P_u.T)
    tq_u = Ti.T @ tq_v - P_u @ S
    return P_u, tq_u, T, S


def prediction_matrices(A, B, N):
    """Literal (tA, tB) of linearMPC.py:397-428 (large: only built on request)."""
    nx, nu = B.shape
    pw = state_powers(A, N)
    tA = pw.reshape((N + 1) * nx, nx)
    tB = np.zeros(((N + 1) * nx, N * nu))
    for d in range(1, N + 1):
        blk = pw[d - 1] @ B
        for k in range(d, N + 1):
            j = k - d
            tB[k * nx:(k + 1) * nx, j * nu:(j + 1) * nu] = blk
    return tA, tB


def extreme_eigs(P, cho=None, seed=0):
    """(lambda_min, lambda_max) of a symmetric positive definite matrix: Lanczos on P for the
    largest and on P^-1 (two triangular solves with the Cholesky factor per apply) for the
    smallest eigenvalue; ~1 s at n = 4480 against ~20 s for a full eigendecomposition."""
    n = P.shape[0]
    if n <= 1024:
        w = np.linalg.eigvalsh(P)
        return float(w[0]), float(w[-1])
    import scipy.sparse.linalg as sla
    v0 = np.random.default_rng(seed).standard_normal(n)
    lmax = float(sla.eigsh(P, k=1, which="LA", v0=v0, tol=1e-3, return_eigenvectors=False)[0])
    cho = scipy.linalg.cho_factor(P, lower=True, check_finite=False) if cho is None else cho
    L = np.tril(cho[0]) if cho[1] else np.triu(cho[0]).T

    def pinv_apply(x):
        y = scipy.linalg.solve_triangular(L, x, lower=True, check_finite=False)
        return scipy.linalg.solve_triangular(L, y, lower=True, trans="T", check_finite=False)

    op = sla.LinearOperator((n, n), matvec=pinv_apply, dtype=np.float64)
    imax = float(sla.eigsh(op, k=1, which="LA", v0=v0, tol=1e-3, maxiter=2000, return_eigenvectors=False)[0])
    return 1.0 / imax, lmax
