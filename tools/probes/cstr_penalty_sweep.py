"""CPU-only: Douglas-Rachford iterations to KKT <= 1e-9 on the CSTR regulator QPs of BASELINE.json configs[1] as a function
of the penalty scale and the relaxation - the FP64 iteration of csrc/qp.cu restated in NumPy for a few hundred QPs
(cold start from the unconstrained law, true KKT residual checked every iteration)."""
import os
import sys
import time

import numpy as np
import scipy.linalg

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from industrial_nnmpc_2021_b200 import condense   # noqa: E402
from industrial_nnmpc_2021_b200.plants import get_cstrs_problem   # noqa: E402
from oracle import linear_mpc as om   # noqa: E402


def main():
    p = get_cstrs_problem()
    aug = om.augmented_matrices_for_regulator(p.A, p.B, p.Q, p.R, p.S)
    K, Pf = condense.dlqr(*aug)
    P, tq = condense.condensed_hessian(*aug, Pf, p.N)
    n, nu = P.shape[0], p.Nu
    ev = np.linalg.eigvalsh(P)
    lmin, lmax = ev[0], ev[-1]
    print(f"n = {n}, cond(P) = {lmax / lmin:.3g}")
    rng = np.random.default_rng(2021)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ots = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs, ulb=p.ulb, uub=p.uub)
    US = np.array([ots.solve(p.setpoints[i][:, None], p.disturbances[j][:, None])[1][:, 0]
                   for i, j in zip(rng.integers(0, p.setpoints.shape[0], 16), rng.integers(0, p.disturbances.shape[0], 16))])
    US = US[rng.integers(0, 16, B)]
    sigma = rng.choice([0.02, 0.1, 0.5, 2.0], B)[:, None]
    dx = sigma * rng.standard_normal((B, p.Nx))
    uprev = p.ulb.T + (p.uub - p.ulb).T * rng.uniform(size=(B, nu))
    X0 = np.hstack([dx, uprev - US])
    LB, UB = np.tile(p.ulb.T - US, (1, p.N)), np.tile(p.uub.T - US, (1, p.N))
    Q = X0 @ tq.T                                   # q = tq x0, per row
    Kunc = -np.linalg.solve(P, tq)
    dP = np.diag(P)
    for rho_scale in (0.25, 0.5, 1.0, 2.0, 4.0):
        rho = 0.5 * np.sqrt(lmin * lmax) * rho_scale * dP / np.exp(np.mean(np.log(dP)))
        Minv = np.linalg.inv(P + np.diag(rho))
        Top, Mq = Minv * rho[None, :], Minv
        C = Q @ Mq.T
        for alpha in (1.6, 1.8, 1.95):
            t0 = time.time()
            V = X0 @ Kunc.T
            its = np.zeros(B, dtype=int)
            live = np.arange(B)
            for it in range(1, 6001):
                Z = np.clip(V[live], LB[live], UB[live])
                W = 2 * Z - V[live]
                X = W @ Top.T - C[live]
                V[live] += alpha * (X - Z)
                if it % 5 == 0 or it < 20:
                    Zn = np.clip(V[live], LB[live], UB[live])
                    G = Zn @ P + Q[live]
                    res = np.max(np.abs(Zn - np.clip(Zn - G, LB[live], UB[live])), axis=1)
                    done = res <= 1e-9
                    its[live[done]] = it
                    live = live[~done]
                    if live.size == 0:
                        break
            its[live] = 6000
            print(f"rho_scale {rho_scale:5.2f} alpha {alpha:4.2f}: iterations mean {its.mean():7.1f} median {np.median(its):6.0f} "
                  f"max {its.max():5d}   ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
