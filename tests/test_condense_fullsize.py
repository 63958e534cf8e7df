"""Full-size pin of the condensed regulator operators (CPU, no GPU): the block recursion of
``condense.condensed_hessian`` at the CDU size (Nxa = 284, N = 140, n = 4480), where the reference's
literal dense form (tQ: 12.8 GB, lib/linearMPC.py:430-474) cannot be built, against a literal
roll-out of the stage costs it condenses (:330-337):

    1/2 u'Pu + (tq x0)'u  ==  V(x0, u) - V(x0, 0)        for arbitrary (x0, u)

The GPU parity test of the headline configuration (tests/test_gpu_cdu_fullsize.py) builds its oracle
on exactly these (P, tq)."""
import numpy as np

from oracle import linear_mpc as om
from industrial_nnmpc_2021_b200 import condense
from industrial_nnmpc_2021_b200.plants import get_cdu_problem


def test_condensed_operators_match_stage_cost_rollout_at_cdu_size():
    p = get_cdu_problem(N=140, with_scenarios=False)
    Aa, Ba, Qa, Ra, Ma = om.augmented_matrices_for_regulator(p.A, p.B, p.Q, p.R, p.S)
    _, Pf = om.dlqr(Aa, Ba, Qa, Ra, Ma)
    P, tq = condense.condensed_hessian(Aa, Ba, Qa, Ra, Ma, Pf, p.N)
    n = p.N * p.Nu
    assert P.shape == (n, n) and tq.shape == (n, p.Nx + p.Nu)
    assert np.array_equal(P, P.T)
    rng = np.random.default_rng(4480)
    for trial in range(4):
        x0 = rng.standard_normal(p.Nx + p.Nu)
        u = rng.uniform(-1.0, 1.0, n) if trial < 3 else np.zeros(n)
        if trial == 2:                       # a sparse sequence: single blocks of P and tq
            u[:] = 0.0
            u[rng.integers(0, n, 5)] = 1.0
        lhs = 0.5 * u @ (P @ u) + (tq @ x0) @ u
        rhs = (om.rollout_cost(Aa, Ba, Qa, Ra, Ma, Pf, p.N, x0, u)
               - om.rollout_cost(Aa, Ba, Qa, Ra, Ma, Pf, p.N, x0, np.zeros(n)))
        scale = max(abs(rhs), 0.5 * u @ (P @ u), 1.0)
        assert abs(lhs - rhs) <= 1e-11 * scale, (trial, lhs, rhs)
    # gradient check on unit vectors: column j of P and row j of tq
    x0 = rng.standard_normal(p.Nx + p.Nu)
    V0 = om.rollout_cost(Aa, Ba, Qa, Ra, Ma, Pf, p.N, x0, np.zeros(n))
    for j in (0, 31, 32, 2239, 4479):
        e = np.zeros(n)
        e[j] = 1.0
        Vp = om.rollout_cost(Aa, Ba, Qa, Ra, Ma, Pf, p.N, x0, e)
        Vm = om.rollout_cost(Aa, Ba, Qa, Ra, Ma, Pf, p.N, x0, -e)
        assert abs((Vp + Vm - 2 * V0) - P[j, j]) <= 1e-10 * max(1.0, P[j, j])          # second difference = P_jj
        assert abs(0.5 * (Vp - Vm) - tq[j] @ x0) <= 1e-10 * max(1.0, abs(tq[j] @ x0))   # first difference = (tq x0)_j
