"""NumPy float64 restatement of the reference's linear-MPC hot path (TEST INFRASTRUCTURE).

Follows /root/reference/lib/linearMPC.py; every function cites the lines it restates.
The formulation is built *literally* (dense ``tA``, ``tB``, block diagonals), exactly as the
reference does, so it is only meant for sizes where that fits (CSTRs N=90; small CDU-like
cases).  The product package builds the same operators by a block recursion; tests compare
the two.  The QP itself is solved exactly (``oracle.qp``), because ``cvxopt`` is absent.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from . import qp as _qp


# ----------------------------------------------------------------------------- helpers
def dlqr(A, B, Q, R, M=None):
    """Discrete LQR with cross term; stage cost x'Qx + 2x'Mu + u'Ru.  linearMPC.py:22-40."""
    if M is None:
        At, Qt, M = A, Q, np.zeros(B.shape)
    else:
        RiMt = scipy.linalg.solve(R, M.T)
        At = A - B @ RiMt
        Qt = Q - M @ RiMt
    Pi = scipy.linalg.solve_discrete_are(At, B, Qt, R)
    K = -scipy.linalg.solve(B.T @ Pi @ B + R, B.T @ Pi @ A + M.T)
    return K, Pi


def c2d(A, B, sample_time):
    """Zero-order-hold discretisation through one matrix exponential.  linearMPC.py:50-64."""
    nx, nu = B.shape
    blk = np.zeros((nx + nu, nx + nu))
    blk[:nx, :nx] = A
    blk[:nx, nx:] = B
    E = scipy.linalg.expm(blk * sample_time)
    return E[:nx, :nx], E[:nx, nx:]


def augmented_matrices_for_regulator(A, B, Q, R, S):
    """Rate-of-change augmentation, state [x; uprev].  linearMPC.py:626-644."""
    nx, nu = B.shape
    Aaug = np.zeros((nx + nu, nx + nu))
    Aaug[:nx, :nx] = A
    Baug = np.vstack([B, np.eye(nu)])
    Qaug = scipy.linalg.block_diag(Q, S)
    Raug = R + S
    Maug = np.vstack([np.zeros((nx, nu)), -S])
    return Aaug, Baug, Qaug, Raug, Maug


# ----------------------------------------------------------------------------- regulator
class DenseQPRegulatorOracle:
    """Literal dense condensing of the regulator QP.  linearMPC.py:321-517.

    min_u 1/2 u'Pu + (tq x0)'u   s.t.  G u <= h(x0)
    """

    def __init__(self, *, A, B, Q, R, M, N, ulb, uub):
        self.A, self.B, self.Q, self.R, self.M, self.N = A, B, Q, R, M, int(N)
        self.ulb, self.uub = ulb, uub
        self.Nx, self.Nu = B.shape
        self.Krep, self.Pf = dlqr(A, B, Q, R, M)                      # :356
        # :366-382 — re-parameterise with u = Kx + v when A is not open-loop stable
        if np.any(np.abs(np.linalg.eigvals(self.A)) >= 1.0):
            K = self.Krep
            self.A = self.A + self.B @ K
            self.Q = self.Q + K.T @ (self.R @ K)
            self.Q = self.Q + self.M @ K + K.T @ self.M.T
            self.M = K.T @ self.R + self.M
            self.reparameterize = True
        else:
            self.reparameterize = False
        self._build()

    def _build(self):
        A, B, N, nx, nu = self.A, self.B, self.N, self.Nx, self.Nu
        # :397-428  tA = [I; A; ...; A^N], tB block lower-triangular Toeplitz
        pw = [np.eye(nx)]
        for _ in range(N):
            pw.append(A @ pw[-1])
        self.tA = np.vstack(pw)
        tB = np.zeros(((N + 1) * nx, N * nu))
        for i in range(1, N + 1):
            for j in range(i):
                tB[i * nx:(i + 1) * nx, j * nu:(j + 1) * nu] = pw[i - j - 1] @ B
        self.tB = tB
        # :430-465 block diagonals
        tQ = scipy.linalg.block_diag(*([self.Q] * N + [self.Pf]))
        tR = scipy.linalg.block_diag(*([self.R] * N))
        tM = np.vstack([scipy.linalg.block_diag(*([self.M] * N)), np.zeros((nx, N * nu))])
        E = np.vstack([np.eye(nu), -np.eye(nu)])
        self.tE = scipy.linalg.block_diag(*([E] * N))
        self.tK = scipy.linalg.block_diag(*([self.Krep] * N)) if self.reparameterize else None
        # :467-474
        self.P = tB.T @ (tQ @ tB) + tR + tB.T @ tM + tM.T @ tB
        self.tq = (tB.T @ tQ + tM.T) @ self.tA
        # :476-482
        if self.reparameterize:
            self.G = self.tE @ (self.tK @ tB[:N * nx, :]) + self.tE
        else:
            self.G = self.tE

    def get_h(self, x0):
        """:484-493 — per-stage [uub; -ulb], shifted by tE tK tA x0 when re-parameterised."""
        te = np.tile(np.vstack([self.uub, -self.ulb]), (self.N, 1))
        if self.reparameterize:
            return te - self.tE @ (self.tK @ (self.tA[:self.N * self.Nx, :] @ x0))
        return te

    def solve(self, x0, return_info=False):
        """:495-512.  Exact solve of the (unique) minimiser; returns useq (n,1)."""
        q = self.tq @ x0
        if self.reparameterize:
            v, info = _qp.solve_general_qp(self.P, q, self.G, self.get_h(x0))
            nN = self.N * self.Nx
            useq = self.tK @ (self.tA[:nN, :] @ x0 + self.tB[:nN, :] @ v) + v
        else:
            lb = np.tile(self.ulb, (self.N, 1))
            ub = np.tile(self.uub, (self.N, 1))
            useq, info = _qp.solve_box_qp(self.P, q, lb, ub)
        return (useq, info) if return_info else useq


class CondensedRegulatorOracle:
    """Box-path regulator on GIVEN condensed operators (P, tq): the same exact solve as
    ``DenseQPRegulatorOracle.solve`` (linearMPC.py:495-512) for sizes where the literal dense
    condensing above does not fit (CDU: tQ is 12.8 GB).  The operators themselves are pinned
    separately: against the literal form at small sizes and, at full size, against a literal
    roll-out of the stage costs (``rollout_cost``)."""

    def __init__(self, *, P, tq, N, Nu, ulb, uub):
        self.P, self.tq, self.N, self.Nu = P, tq, int(N), int(Nu)
        self.ulb, self.uub = ulb, uub
        self.reparameterize = False
        self._box = _qp.BoxQP(P)

    def solve(self, x0, return_info=False):
        q = self.tq @ x0
        useq, info = self._box.solve(q, np.tile(self.ulb, (self.N, 1)), np.tile(self.uub, (self.N, 1)))
        useq = useq.reshape(-1, 1)
        return (useq, info) if return_info else useq


def rollout_cost(A, B, Q, R, M, Pf, N, x0, useq):
    """1/2 sum_k (x_k'Q x_k + u_k'R u_k + 2 x_k'M u_k) + 1/2 x_N'Pf x_N along x+ = Ax + Bu: the objective
    the reference condenses into 1/2 u'Pu + (tq x0)'u + const (linearMPC.py:330-337, :467-474)."""
    nu = B.shape[1]
    x = np.asarray(x0, float).reshape(-1)
    u = np.asarray(useq, float).reshape(N, nu)
    J = 0.0
    for k in range(N):
        J += x @ (Q @ x) + u[k] @ (R @ u[k]) + 2.0 * (x @ (M @ u[k]))
        x = A @ x + B @ u[k]
    return 0.5 * (J + x @ (Pf @ x))


# ----------------------------------------------------------------------------- target selector
class TargetSelectorOracle:
    """Steady-state target QP in (xs, us).  linearMPC.py:178-319 (input-bound branch :249-251, output-bound
    branch :242-248 when both ylb and yub are given)."""

    def __init__(self, *, A, B, C, H, Bd, Cd, usp, Rs, Qs, ulb, uub, ylb=None, yub=None):
        self.A, self.B, self.C, self.H, self.Bd, self.Cd = A, B, C, H, Bd, Cd
        self.usp, self.Rs, self.Qs, self.ulb, self.uub = usp, Rs, Qs, ulb, uub
        self.ylb, self.yub = ylb, yub
        self.Nx, self.Nu = B.shape
        self.Ny, self.Nd, self.Nz = C.shape[0], Bd.shape[1], H.shape[0]
        nx, nu, ny, nz = self.Nx, self.Nu, self.Ny, self.Nz
        E = np.vstack([np.eye(nu), -np.eye(nu)])
        self.F = np.vstack([np.eye(ny), -np.eye(ny)])                     # :240
        if ylb is not None and yub is not None:                           # :242-248
            self.G = np.block([[self.F @ C, np.zeros((2 * ny, nu))], [np.zeros((2 * nu, nx)), E]])
            self.f = np.vstack([yub, -ylb])
            self.e = np.vstack([uub, -ulb])
            self.h = None
        else:
            self.G = np.hstack([np.zeros((2 * nu, nx)), E])              # :250
            self.h = np.vstack([uub, -ulb])                               # :251
        self.tA = np.block([[np.eye(nx) - A, -B], [H @ C, np.zeros((nz, nu))]])     # :254-260
        self.tb = np.block([[np.zeros((nx, ny)), Bd], [H, -(H @ Cd)]])              # :261-267
        self.P = scipy.linalg.block_diag(C.T @ (Qs @ C), Rs)             # :270-274

    def changing(self, ysp, dhats):
        """:276-296."""
        q = np.vstack([-(self.C.T @ (self.Qs @ (ysp - self.Cd @ dhats))), -(self.Rs @ self.usp)])
        b = self.tb @ np.vstack([ysp, dhats])
        if self.h is None:                                                # :284-288
            h = np.vstack([self.f - self.F @ (self.Cd @ dhats), self.e])
        else:
            h = self.h
        return q, h, b

    def solve(self, ysp, dhats, return_info=False):
        """:298-311 — returns (xs, us)."""
        q, h, b = self.changing(ysp, dhats)
        nx = self.Nx
        if self.h is None:
            # general inequalities next to the equalities: eliminate tA w = b through a null-space basis, then the
            # inequality QP of oracle.qp (interior point + active-set polish); raises when no feasible point exists
            wp = np.linalg.lstsq(self.tA, b, rcond=None)[0]
            Z = scipy.linalg.null_space(self.tA)
            Pr = Z.T @ self.P @ Z
            y, info = _qp.solve_general_qp(0.5 * (Pr + Pr.T), Z.T @ (q + self.P @ wp), self.G @ Z, h - self.G @ wp)
            w = wp + Z @ y
            if not np.all(np.isfinite(w)) or not np.max(self.G @ w - h) <= 1e-7 * (1.0 + np.abs(h).max()):
                raise ValueError("TargetSelectorOracle: the output-constrained target problem is infeasible")
            xs, us = w[:nx], w[nx:]
            return ((xs, us), info) if return_info else (xs, us)
        lb = np.vstack([np.full((nx, 1), -np.inf), self.ulb])
        ub = np.vstack([np.full((nx, 1), np.inf), self.uub])
        w, info = _qp.solve_eq_box_qp(self.P, q, self.tA, b, lb, ub)
        xs, us = w[:nx], w[nx:]
        return ((xs, us), info) if return_info else (xs, us)


# ----------------------------------------------------------------------------- controller glue
def get_control_sequence(regulator, x, uprev, xs, us, ulb, uub):
    """linearMPC.py:682-689 — bounds shifted by us, x0 in deviation variables, us added back."""
    regulator.ulb = ulb - us
    regulator.uub = uub - us
    x0 = np.vstack([x - xs, uprev - us])
    return regulator.solve(x0) + np.tile(us, (regulator.N, 1))


def updated_average_stage_cost(x, uprev, xs, us, u, Qaug, Raug, Maug, avg, time_index):
    """linearMPC.py:691-701 — running mean of x'Qx + u'Ru + 2x'Mu (no 1/2 factor)."""
    xa = np.vstack([x - xs, uprev - us])
    du = u - us
    ell = xa.T @ (Qaug @ xa) + du.T @ (Raug @ du) + xa.T @ (Maug @ du) + du.T @ (Maug.T @ xa)
    return (avg * (time_index - 1) + ell) / time_index


def setup_regulator(A, B, Q, R, S, N, ulb, uub):
    """linearMPC.py:596-604."""
    Aa, Ba, Qa, Ra, Ma = augmented_matrices_for_regulator(A, B, Q, R, S)
    return DenseQPRegulatorOracle(A=Aa, B=Ba, Q=Qa, R=Ra, M=Ma, N=N, ulb=ulb, uub=uub)


def split_scenarios(setpoints, disturbances, num_processes):
    """linearMPC.py:786-801 — equal chunks, remainder rows dropped."""
    L = int(setpoints.shape[0] / num_processes)
    return ([setpoints[i * L:(i + 1) * L] for i in range(num_processes)],
            [disturbances[i * L:(i + 1) * L] for i in range(num_processes)])


def simulate_offline(*, x0, uprev0, A, B, Bd, regulator, ulb, uub, target_selector,
                     setpoints, disturbances):
    """The closed loop of linearMPC.py:827-880 for one chunk.

    Returns dict(x, uprev, xs, us, u) with shapes (L,Nx),(L,Nu),(L,Nx),(L,Nu),(L,Nu);
    row t holds the state *before* step t (:868-872).
    """
    nu = B.shape[1]
    xt, upt = x0, uprev0
    rows = dict(x=[], uprev=[], xs=[], us=[], u=[])
    for t in range(setpoints.shape[0]):
        ysp = setpoints[t][:, None]
        d = disturbances[t][:, None]
        xs, us = target_selector.solve(ysp, d)
        useq = get_control_sequence(regulator, xt, upt, xs, us, ulb, uub)
        ut = useq[:nu]
        for k, v in zip(("x", "uprev", "xs", "us", "u"), (xt, upt, xs, us, ut)):
            rows[k].append(v[:, 0])
        xt = A @ xt + B @ ut + Bd @ d
        upt = ut
    return {k: np.asarray(v) for k, v in rows.items()}


# ----------------------------------------------------------------------------- online loop
def dlqe(A, C, Q, R):
    """Steady-state Kalman gain.  linearMPC.py:42-48."""
    P = scipy.linalg.solve_discrete_are(A.T, C.T, Q, R)
    L = scipy.linalg.solve(C @ P @ C.T + R, C @ P).T
    return L, P


def augmented_matrices_for_filter(A, B, C, Bd, Cd, Qwx, Qwd):
    """Integrating-disturbance model [x; d].  linearMPC.py:606-624."""
    nx, nu, nd = A.shape[0], B.shape[1], Bd.shape[1]
    Aaug = np.block([[A, Bd], [np.zeros((nd, nx)), np.eye(nd)]])
    Baug = np.vstack([B, np.zeros((nd, nu))])
    Caug = np.hstack([C, Cd])
    return Aaug, Baug, Caug, scipy.linalg.block_diag(Qwx, Qwd)


class OnlineControllerOracle:
    """LinearMPCController.control_law (linearMPC.py:646-669): Kalman filter (:133-176) -> target selector
    -> regulator -> running average stage cost (:691-701), one measurement at a time."""

    def __init__(self, *, A, B, C, H, Qwx, Qwd, Rv, xprior, dprior, Rs, Qs, Bd, Cd, usp, uprev, Q, R, S, ulb, uub, N,
                 regulator=None):
        self.Nx, self.Nu = B.shape
        self.ulb, self.uub = ulb, uub
        self.Af, self.Bf, self.Cf, Qw = augmented_matrices_for_filter(A, B, C, Bd, Cd, Qwx, Qwd)
        self.L, _ = dlqe(self.Af, self.Cf, Qw, Rv)
        self.xhat = np.vstack([xprior, dprior])
        self.target_selector = TargetSelectorOracle(A=A, B=B, C=C, H=H, Bd=Bd, Cd=Cd, usp=usp, Rs=Rs, Qs=Qs,
                                                    ulb=ulb, uub=uub)
        self.regulator = regulator if regulator is not None else setup_regulator(A, B, Q, R, S, N, ulb, uub)
        _, _, self.Qaug, self.Raug, self.Maug = augmented_matrices_for_regulator(A, B, Q, R, S)
        self.uprev = uprev
        self.average_stage_costs = [np.zeros((1, 1))]
        self.xhats, self.targets = [], []

    def control_law(self, ysp, y):
        pred = self.Af @ self.xhat + self.Bf @ self.uprev                       # :163-165
        self.xhat = pred + self.L @ (y - self.Cf @ pred)
        xhat, dhat = self.xhat[:self.Nx], self.xhat[self.Nx:]
        xs, us = self.target_selector.solve(ysp, dhat)
        u = self.control_input(xhat, xs, us)
        self.average_stage_costs.append(updated_average_stage_cost(
            xhat, self.uprev, xs, us, u, self.Qaug, self.Raug, self.Maug, self.average_stage_costs[-1],
            len(self.average_stage_costs)))
        self.uprev = u
        self.xhats.append(xhat)
        self.targets.append((xs, us))
        return u


    def control_input(self, xhat, xs, us):
        """The MPC law (:658-661); NeuralNetworkController / SatDlqrController replace exactly this step
        (controller_evaluation.py:849-852, :985-987)."""
        useq = get_control_sequence(self.regulator, xhat, self.uprev, xs, us, self.ulb, self.uub)
        return useq[:self.Nu]


class OnlineNNControllerOracle(OnlineControllerOracle):
    """NeuralNetworkController.control_law (controller_evaluation.py:841-862)."""

    def __init__(self, *, regulator_weights, xscale, nnwithuprev, **kw):
        super().__init__(regulator=object(), **kw)
        self.weights, self.xscale, self.nnwithuprev = regulator_weights, np.asarray(xscale, float).reshape(-1, 1), nnwithuprev

    def control_input(self, xhat, xs, us):
        from . import nn as _nn
        return _nn.control_input(self.weights, xhat, self.uprev, xs, us, self.nnwithuprev, self.xscale, self.ulb, self.uub)


class OnlineSatDlqrControllerOracle(OnlineControllerOracle):
    """SatDlqrController.control_law (controller_evaluation.py:975-993)."""

    def __init__(self, **kw):
        super().__init__(regulator=object(), **kw)
        Aa, Ba, Qa, Ra, Ma = augmented_matrices_for_regulator(kw["A"], kw["B"], kw["Q"], kw["R"], kw["S"])
        self.Kaug, _ = dlqr(Aa, Ba, Qa, Ra, Ma)

    def control_input(self, xhat, xs, us):
        u = self.Kaug @ np.vstack([xhat - xs, self.uprev - us]) + us
        u = np.where(u > self.uub, self.uub, u)
        return np.where(u < self.ulb, self.ulb, u)


def online_simulation(plant_step, y0, controller, setpoints, disturbances, Nsim):
    """linearMPC.py:703-718 with the plant given as a callable (u, p) -> y."""
    y = y0
    us = []
    for i in range(min(Nsim, setpoints.shape[0])):
        u = controller.control_law(setpoints[i][:, None], y)
        y = plant_step(u, disturbances[i][:, None])
        us.append(u[:, 0])
    return np.asarray(us)
