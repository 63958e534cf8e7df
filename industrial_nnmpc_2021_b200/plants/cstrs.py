"""CSTRs-in-series-with-flash example: plant, linear model, MPC tuning and offline scenarios.

Numerical restatement of /root/reference/cstrs_parameters.py without casadi/mpctools:

* ODE right-hand side                      cstrs_parameters.py:24-102
* parameters, bounds, scaling              cstrs_parameters.py:110-204
* rectified steady state (7200 x 10 s at zero deviation input)      :206-223
  -> here: stiff integration over the same 72 000 s followed by a Newton polish of f(x)=0
* linearisation at zero deviation + ZOH discretisation, Delta = 10 s :225-246
  -> here: complex-step Jacobians (exact to round-off, like casadi AD) and the same
     augmented-matrix exponential as ``c2d`` (lib/linearMPC.py:50-64)
* MPC tuning                               cstrs_parameters.py:263-312
* offline scenarios (PRBS-like signals)    cstrs_parameters.py:314-351, :414-419
"""
from __future__ import annotations

import functools
import numpy as np
import scipy.integrate
import scipy.linalg
import scipy.optimize

from .problem import MPCProblem
from ..controller_evaluation import sample_prbs_like

Z_INDICES = (0, 3, 4, 7, 8, 11)
UNEXP_Z_INDICES = (4,)
EXP_DIST_INDICES = (0, 1, 2, 3, 4)


def cstrs_parameters():
    """Physical constants, nominal point, bounds and scalings (cstrs_parameters.py:110-204)."""
    par = dict(alphaA=3.5, alphaB=1.1, alphaC=0.5, pho=50., Cp=3., Ar=0.3, Am=2., Ab=4.,
               kr=2.5, km=2.5, kb=1.5, delH1=-40., delH2=-50., EbyR=150., k1star=4e-4,
               k2star=1.8e-6, Td=313., Nx=12, Nu=6, Np=5, Ny=12, sample_time=10.)
    par["xs"] = np.array([178.56, 1, 0, 313, 190.07, 1, 0, 313, 5.17, 1, 0, 313.])
    par["us"] = np.array([2., 0., 1., 0., 30., 0.])
    par["ps"] = np.array([0.8, 0.1, 0.8, 0.1, 313.])
    ulb = np.array([-0.5, -500., -0.5, -500., -0.5, -500.])
    uub = -ulb
    ylb = np.array([-5., 0., 0., -10., -5., 0., 0., -3., -1., 0., 0., -10.])
    yub = np.array([5., 1., 1., 10., 5., 1., 1., 3., 1., 1., 1., 10.])
    plb = np.array([-0.1, -0.1, -0.1, -0.1, -8.])
    pub = np.array([0.05, 0.05, 0.05, 0.05, 8.])
    par["uscale"] = 0.5 * (uub - ulb)
    par["pscale"] = 0.5 * (pub - plb)
    par["yscale"] = 0.5 * (yub - ylb)
    par["lb"] = dict(u=ulb / par["uscale"], y=ylb / par["yscale"], p=plb / par["pscale"])
    par["ub"] = dict(u=uub / par["uscale"], y=yub / par["yscale"], p=pub / par["pscale"])
    par["C"] = np.eye(12)
    return par


def cstrs_ode(x, u, p, par):
    """dx/dt in deviation variables; u, p are scaled deviations (cstrs_parameters.py:24-102).

    Written with operations that stay analytic for complex arguments (complex-step Jacobian).
    """
    X = x + par["xs"]
    Hr, xAr, xBr, Tr, Hm, xAm, xBm, Tm, Hb, xAb, xBb, Tb = X
    F0, Qr, F1, Qm, D, Qb = u * par["uscale"] + par["us"]
    xA0, xB0, xA1, xB1, T0 = p * par["pscale"] + par["ps"]
    aA, aB, aC = par["alphaA"], par["alphaB"], par["alphaC"]
    rho, Cp, Td = par["pho"], par["Cp"], par["Td"]
    dH1, dH2 = par["delH1"], par["delH2"]
    # flash vapour composition
    den = aA * xAb + aB * xBb + aC * (1 - xAb - xBb)
    xAd, xBd = aA * xAb / den, aB * xBb / den
    # outlet flows
    Fr, Fm, Fb = par["kr"] * np.sqrt(Hr), par["km"] * np.sqrt(Hm), par["kb"] * np.sqrt(Hb)
    Fp = 0.01 * D
    # Arrhenius rates
    k1r, k2r = par["k1star"] * np.exp(-par["EbyR"] / Tr), par["k2star"] * np.exp(-par["EbyR"] / Tr)
    k1m, k2m = par["k1star"] * np.exp(-par["EbyR"] / Tm), par["k2star"] * np.exp(-par["EbyR"] / Tm)
    mr, mm, mb = rho * par["Ar"], rho * par["Am"], rho * par["Ab"]
    f = [
        (F0 + D - Fr) / mr,
        (F0 * (xA0 - xAr) + D * (xAd - xAr)) / (mr * Hr) - k1r * xAr,
        (F0 * (xB0 - xBr) + D * (xBd - xBr)) / (mr * Hr) + k1r * xAr - k2r * xBr,
        (F0 * (T0 - Tr) + D * (Td - Tr)) / (mr * Hr) - (k1r * xAr * dH1 + k2r * xBr * dH2) / Cp
        + Qr / (mr * Cp * Hr),
        (Fr + F1 - Fm) / mm,
        (Fr * (xAr - xAm) + F1 * (xA1 - xAm)) / (mm * Hm) - k1m * xAm,
        (Fr * (xBr - xBm) + F1 * (xB1 - xBm)) / (mm * Hm) + k1m * xAm - k2m * xBm,
        (Fr * (Tr - Tm) + F1 * (T0 - Tm)) / (mm * Hm) - (k1m * xAm * dH1 + k2m * xBm * dH2) / Cp
        + Qm / (mm * Cp * Hm),
        (Fm - Fb - D - Fp) / mb,
        (Fm * (xAm - xAb) - (D + Fp) * (xAd - xAb)) / (mb * Hb),
        (Fm * (xBm - xBb) - (D + Fp) * (xBd - xBb)) / (mb * Hb),
        (Fm * (Tm - Tb)) / (mb * Hb) + Qb / (mb * Cp * Hb),
    ]
    return np.array(f)


def rectified_steady_state(par):
    """Steady state reached from the nominal guess at zero deviation inputs (:206-223)."""
    u0, p0 = np.zeros(6), np.zeros(5)
    sol = scipy.integrate.solve_ivp(lambda t, x: cstrs_ode(x, u0, p0, par), (0., 7200 * 10.),
                                    np.zeros(12), method="Radau", rtol=1e-11, atol=1e-12)
    x = sol.y[:, -1]
    x = scipy.optimize.fsolve(lambda x: cstrs_ode(x, u0, p0, par), x, xtol=1e-14)
    return par["xs"] + x


def _complex_step_jac(fun, z0, h=1e-30):
    n = z0.size
    cols = []
    for i in range(n):
        z = z0.astype(complex)
        z[i] += 1j * h
        cols.append(np.imag(fun(z)) / h)
    return np.stack(cols, axis=1)


def linearised_model(par):
    """(A, B, C, Bp) of the discrete linear model at zero deviation (:225-246)."""
    z = np.zeros
    Ac = _complex_step_jac(lambda x: cstrs_ode(x, z(6), z(5), par), z(12))
    Bc = _complex_step_jac(lambda u: cstrs_ode(z(12), u, z(5), par), z(6))
    Bpc = _complex_step_jac(lambda p: cstrs_ode(z(12), z(6), p, par), z(5))
    blk = np.zeros((23, 23))
    blk[:12, :12], blk[:12, 12:18], blk[:12, 18:] = Ac, Bc, Bpc
    E = scipy.linalg.expm(blk * par["sample_time"])
    C = np.diag(1.0 / par["yscale"]) @ par["C"]
    return E[:12, :12], E[:12, 12:18], C, E[:12, 18:]


@functools.lru_cache(maxsize=4)
def _cached_model():
    par = cstrs_parameters()
    par["xs"] = rectified_steady_state(par)
    return par, linearised_model(par)


def get_cstrs_problem(*, N=90, Nsim=150000, seed=1, conservative_factor=1.02,
                      with_scenarios=True) -> MPCProblem:
    """MPC problem of cstrs_parameters.py:263-351 with the code's tuning (N = 90)."""
    par, (A, B, C, Bp) = _cached_model()
    Nx, Nu, Ny = 12, 6, 12
    Bd = Bp[:, EXP_DIST_INDICES]
    Nd = Bd.shape[1]
    Qs = np.zeros((Ny, Ny))
    Qs[Z_INDICES, Z_INDICES] = 1.0
    prob = MPCProblem(
        name="cstrs", A=A.copy(), B=B.copy(), C=C.copy(), H=np.zeros((0, Ny)), Bd=Bd,
        Cd=np.zeros((Ny, Nd)), Q=1e3 * (C.T @ C), R=0.1 * np.eye(Nu), S=0.1 * np.eye(Nu), N=N,
        Rs=np.zeros((Nu, Nu)), Qs=Qs, usp=np.zeros((Nu, 1)),
        ulb=par["lb"]["u"][:, None].copy(), uub=par["ub"]["u"][:, None].copy(),
        xprior=np.zeros((Nx, 1)), uprev=np.zeros((Nu, 1)),
        extra=dict(parameters=par, z_indices=Z_INDICES))
    if with_scenarios:
        # cstrs_parameters.py:314-351
        ylb, yub = par["lb"]["y"] * conservative_factor, par["ub"]["y"] * conservative_factor
        plb, pub = par["lb"]["p"] * conservative_factor, par["ub"]["p"] * conservative_factor
        sp_y = sample_prbs_like(num_change=1250, num_steps=Nsim, lb=ylb, ub=yub,
                                mean_change=120, sigma_change=2, seed=seed)
        sp = np.zeros((Nsim, Ny))
        sp[:, Z_INDICES] = sp_y[:, Z_INDICES]
        sp[:, UNEXP_Z_INDICES] = 0.0
        dist = sample_prbs_like(num_change=2500, num_steps=Nsim, lb=plb, ub=pub,
                                mean_change=60, sigma_change=5, seed=seed + 1)
        prob.setpoints, prob.disturbances = sp, dist[:, EXP_DIST_INDICES]
    return prob
