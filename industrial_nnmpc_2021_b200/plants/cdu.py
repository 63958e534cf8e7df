"""Crude-distillation-unit (CDU) example: synthetic stand-in plant + the reference's tuning.

The reference loads the plant from ``CDU_Model.mat`` (/root/reference/cdu_parameters.py:200),
which is NOT shipped, so the true matrices are unavailable.  This module builds a documented,
seeded synthetic model with the reference's dimensions (252 states, 32 inputs, 90 outputs,
5 disturbances = input channels (0, 6, 23, 30, 31), cdu_parameters.py:198) and applies the
reference's own scaling conventions, tuning and scenario generators on top of it:

* scaled bounds +-1 on u and y                       cdu_parameters.py:22-52
* Bd = B[:, dist_indices], Cd = 0, H = 0 x Ny        cdu_parameters.py:49, :78-83
* Rs = 1e-6 I, Qs = diag(1e-16 I_86, I_4)            cdu_parameters.py:94-96
* Q = 2 C'C, R = 0.1 I, S = 0, N = 140               cdu_parameters.py:99-102
* offline scenarios: 894 / 1788 changes, mean hold 400 / 200, sigma 1, seeds 1 / 2,
  conservative factor 1.05, dist_scaling [5,20,20,20,20], Nsim = 357 600
                                                      cdu_parameters.py:115-143, :199, :211

Synthetic plant (seed 252): A is block diagonal with 84 first-order lags and 84 lightly
oscillatory second-order blocks, time constants log-uniform in [5, 120] min at a 1 min sample
time (|lambda| in [0.82, 0.992], open-loop stable like the reference's box-constrained path
requires); B and C are 20 %-dense Gaussian, B rows carry (1 - pole) so every state has O(1) DC
gain, C rows are normalised so that each output's steady-state gain row has 2-norm GAIN_NORM.
The five disturbance channels carry gain ``DIST_GAIN/dist_scaling`` so that a full-range
disturbance (+-5..20 in scaled units, as in the reference) moves the outputs about twice as much
as one input does.  GAIN_NORM = 0.7 / DIST_GAIN = 2 were picked so that the closed loop is
moderately constrained (probe over 240 closed-loop samples: 45 % of the QPs have active bounds,
26 active of 4480 on average, up to 201; cond(P) = 72) - comparable to the reference's own CSTR
example (73 % of QPs constrained) rather than an all-LQR regime.
"""
from __future__ import annotations

import functools
import numpy as np
import scipy.linalg

from .problem import MPCProblem
from ..controller_evaluation import sample_prbs_like

DIST_INDICES = (0, 6, 23, 30, 31)
DIST_SCALING = np.array([[5., 20., 20., 20., 20.]])
NZ = 4
GAIN_NORM = 0.7      # 2-norm of each output's steady-state gain row
DIST_GAIN = 2.0      # disturbance-channel gain relative to 1/dist_scaling


def _dist_indices(Nu):
    """Reference channels for the full model; evenly spread ones for reduced test models."""
    if Nu >= 32:
        return DIST_INDICES
    return tuple(sorted(set(int(round(i * (Nu - 1) / 4)) for i in range(5))))


@functools.lru_cache(maxsize=4)
def synthetic_cdu_model(Nx=252, Nu=32, Ny=90, seed=252, density=0.2, gain_norm=GAIN_NORM,
                        dist_gain=DIST_GAIN):
    dist_idx = _dist_indices(Nu)
    rng = np.random.default_rng(seed)
    n1 = Nx // 3                       # first-order lags
    n2 = (Nx - n1) // 2                # second-order blocks
    n1 = Nx - 2 * n2
    tau = np.exp(rng.uniform(np.log(5.0), np.log(120.0), size=n1 + n2))
    pole = np.exp(-1.0 / tau)
    theta = rng.uniform(0.0, 0.12, size=n2)
    blocks, dc = [], []
    for i in range(n1):
        blocks.append(np.array([[pole[i]]]))
        dc.append(1.0 - pole[i])
    for j in range(n2):
        r, th = pole[n1 + j], theta[j]
        blocks.append(r * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]))
        dc += [1.0 - r, 1.0 - r]
    A = scipy.linalg.block_diag(*blocks)
    perm = rng.permutation(Nx)         # mix first/second-order states
    A = A[np.ix_(perm, perm)]
    dc = np.asarray(dc)[perm]
    B = rng.standard_normal((Nx, Nu)) * (rng.random((Nx, Nu)) < density)
    B = B * dc[:, None]
    C = rng.standard_normal((Ny, Nx)) * (rng.random((Ny, Nx)) < density)
    B[:, dist_idx] = B[:, dist_idx] * (dist_gain / DIST_SCALING[:, :len(dist_idx)])
    Gss = C @ np.linalg.solve(np.eye(Nx) - A, B)
    C = C * (gain_norm / np.linalg.norm(Gss, axis=1, keepdims=True))
    for M in (A, B, C):
        M.setflags(write=False)
    return A, B, C


def get_cdu_problem(*, N=140, Nsim=357600, seed=1, conservative_factor=1.05,
                    with_scenarios=True, Nx=252, Nu=32, Ny=90, model_seed=252,
                    gain_norm=GAIN_NORM, dist_gain=DIST_GAIN) -> MPCProblem:
    """CDU MPC problem with the reference's tuning on the synthetic stand-in plant.

    Smaller (Nx, Nu, Ny) build a reduced model of the same family for quick parity tests.
    """
    A, B, C = synthetic_cdu_model(Nx, Nu, Ny, model_seed, 0.2, gain_norm, dist_gain)
    dist_idx = _dist_indices(Nu)
    dist_scaling = DIST_SCALING[:, :len(dist_idx)]
    Bd = np.take(B, dist_idx, axis=1)
    Nd = Bd.shape[1]
    lb = dict(u=-np.ones((Nu, 1)), y=-np.ones((Ny, 1)))
    ub = dict(u=np.ones((Nu, 1)), y=np.ones((Ny, 1)))
    prob = MPCProblem(
        name="cdu", A=A.copy(), B=B.copy(), C=C.copy(), H=np.zeros((0, Ny)), Bd=Bd,
        Cd=np.zeros((Ny, Nd)), Q=2.0 * (C.T @ C), R=0.1 * np.eye(Nu), S=0.0 * np.eye(Nu), N=N,
        Rs=1e-6 * np.eye(Nu), Qs=scipy.linalg.block_diag(1e-16 * np.eye(Ny - NZ), np.eye(NZ)),
        usp=np.zeros((Nu, 1)), ulb=lb["u"], uub=ub["u"], xprior=np.zeros((Nx, 1)),
        uprev=np.zeros((Nu, 1)), extra=dict(dist_indices=dist_idx, synthetic=True))
    if with_scenarios:
        Hsel = np.hstack([np.zeros((NZ, Ny - NZ)), np.eye(NZ)])
        sp_lb = (Hsel @ lb["y"]) * conservative_factor
        sp_ub = (Hsel @ ub["y"]) * conservative_factor
        d_lb = (np.take(lb["u"], dist_idx) * dist_scaling * conservative_factor).reshape(-1, 1)
        d_ub = (np.take(ub["u"], dist_idx) * dist_scaling * conservative_factor).reshape(-1, 1)
        # the reference's 894 / 1788 change points over 357 600 steps (mean hold 400 / 200); other
        # Nsim keep the same hold statistics
        nch_sp = 894 if Nsim == 357600 else max(2, Nsim // 400)
        nch_d = 1788 if Nsim == 357600 else max(2, Nsim // 200)
        sp = sample_prbs_like(num_change=nch_sp, num_steps=Nsim, lb=sp_lb, ub=sp_ub,
                              mean_change=400, sigma_change=1, seed=seed)
        prob.setpoints = np.hstack([np.zeros((Nsim, Ny - NZ)), sp])
        prob.disturbances = sample_prbs_like(num_change=nch_d, num_steps=Nsim, lb=d_lb, ub=d_ub,
                                             mean_change=200, sigma_change=1, seed=seed + 1)
    return prob
