"""CPU-only NumPy restatement of k_ts_general (csrc/ts.cu): the range-space dual active-set method for the
output-constrained target problem, checked against the oracle on random plants.  Same operators (Hinv, Abar, AH, Mbar),
same step rules and tolerances as the kernel; used to settle signs and the drop rule before GPU time is spent."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import linear_mpc as om   # noqa: E402


def dual_active_set(Hinv, Abar, AH, Mbar, f, lo, hi, nu):
    mb = Abar.shape[0]
    u = -Hinv @ f
    W, sg, lam = [], [], []
    maxit = 8 * (nu + 8) + 2 * (mb - nu)
    for it in range(maxit):
        val = Abar @ u
        mag = np.abs(Abar) @ np.abs(u)
        vu, vl = val - hi, lo - val
        v = np.maximum(vu, vl)
        tol = 1e-11 * (1.0 + mag + np.abs(np.where(vu >= vl, hi, lo)))
        v = np.where(v > tol, v, 0.0)
        ip = int(np.argmax(v))
        if not v[ip] > 0.0:
            break
        sp = 1 if vu[ip] >= vl[ip] else -1
        vp, lam_p = v[ip], 0.0
        for inner in range(2 * nu + 3):
            nW = len(W)
            d = np.array([sg[j] * sp * Mbar[W[j], ip] for j in range(nW)])
            S = np.array([[sg[j] * sg[k] * Mbar[W[j], W[k]] for k in range(nW)] for j in range(nW)]).reshape(nW, nW)
            r = np.linalg.solve(S, d) if nW else np.zeros(0)
            z = sp * AH[ip] - sum((r[j] * sg[j] * AH[W[j]] for j in range(nW)), np.zeros(nu))
            apz = Mbar[ip, ip] - d @ r
            t1l = [lam[j] / r[j] if r[j] > 1e-13 * (1.0 + abs(d[j])) else np.inf for j in range(nW)]
            t1 = min(t1l) if t1l else np.inf
            t2 = vp / apz if apz > 1e-12 * Mbar[ip, ip] else np.inf
            if not np.isfinite(t1) and not np.isfinite(t2):
                return None, -it - 1
            t = min(t1, t2)
            u = u - t * z
            lam = [lam[j] - t * r[j] for j in range(nW)]
            lam_p += t
            vp -= t * apz
            if t2 <= t1:
                W.append(ip); sg.append(sp); lam.append(lam_p)
                break
            jb = int(np.argmin(t1l))
            del W[jb], sg[jb], lam[jb]
        else:
            return None, -it - 1
    else:
        return None, -maxit - 1
    if W:
        nW = len(W)
        S = np.array([[sg[j] * sg[k] * Mbar[W[j], W[k]] for k in range(nW)] for j in range(nW)])
        rhs = -np.array([(hi[W[j]] if sg[j] > 0 else -lo[W[j]]) + sg[j] * (AH[W[j]] @ f) for j in range(nW)])
        lm = np.linalg.solve(S, rhs)
        u = -Hinv @ f - sum((lm[j] * sg[j] * AH[W[j]] for j in range(nW)), np.zeros(nu))
    return u, it


def main():
    rng = np.random.default_rng(1)
    worst, n_act_tot, n_inf = 0.0, 0, 0
    for trial in range(200):
        nx, nu, ny, nd = int(rng.integers(4, 12)), int(rng.integers(1, 6)), int(rng.integers(1, 7)), 2
        A = np.diag(rng.uniform(0.2, 0.9, nx)) + 0.05 * rng.standard_normal((nx, nx))
        B, C = rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx))
        Bd, Cd = rng.standard_normal((nx, nd)), 0.1 * rng.standard_normal((ny, nd))
        Rs, Qs = 10.0 ** rng.uniform(-4, 0) * np.eye(nu), np.eye(ny)
        ulb, uub = -np.ones((nu, 1)), np.ones((nu, 1))
        kw = dict(A=A, B=B, C=C, H=np.zeros((0, ny)), Bd=Bd, Cd=Cd, usp=0.1 * rng.standard_normal((nu, 1)), Rs=Rs, Qs=Qs,
                  ulb=ulb, uub=uub)
        ysp, d = rng.standard_normal((ny, 1)), 0.2 * rng.standard_normal((nd, 1))
        xs0, us0 = om.TargetSelectorOracle(**kw).solve(ysp, d)
        y0 = C @ xs0 + Cd @ d
        width = 10.0 ** rng.uniform(-2, 0.5)
        ylb, yub = y0 - width * rng.uniform(0, 1, (ny, 1)) + 0.3 * width, y0 + width * rng.uniform(0, 1, (ny, 1)) - 0.3 * width
        ylb, yub = np.minimum(ylb, yub), np.maximum(ylb, yub)
        # reduced operators exactly as linearMPC.TargetSelector builds them
        ImA = np.eye(nx) - A
        Gx, Gd = np.linalg.solve(ImA, B), np.linalg.solve(ImA, Bd)
        CG = C @ Gx
        Ht = CG.T @ Qs @ CG + Rs
        Ht = 0.5 * (Ht + Ht.T)
        f = (-(Qs @ CG).T @ ysp + (Qs @ CG).T @ ((C @ Gd + Cd) @ d) - Rs @ kw["usp"]).ravel()
        Abar = np.vstack([CG, np.eye(nu)])
        Hinv = np.linalg.inv(Ht)
        AH = Abar @ Hinv
        Mbar = AH @ Abar.T
        r = ((C @ Gd + Cd) @ d).ravel()
        lo = np.concatenate([ylb.ravel() - r, ulb.ravel()])
        hi = np.concatenate([yub.ravel() - r, uub.ravel()])
        u, it = dual_active_set(Hinv, Abar, AH, Mbar, f, lo, hi, nu)
        # independent feasibility verdict: phase-1 LP  min s  s.t.  lo - s <= Abar u <= hi + s, s >= 0
        from scipy.optimize import linprog
        lp = linprog(np.r_[np.zeros(nu), 1.0], A_ub=np.block([[Abar, -np.ones((len(lo), 1))], [-Abar, -np.ones((len(lo), 1))]]),
                     b_ub=np.r_[hi, -lo], bounds=[(None, None)] * nu + [(0, None)], method="highs")
        feasible = lp.status == 0 and lp.x[-1] <= 1e-9
        if not feasible:
            assert u is None, f"trial {trial}: the kernel restatement returned a point for an infeasible problem"
            n_inf += 1
            continue
        assert u is not None, f"trial {trial}: dual active set failed on a feasible problem (it={it})"
        (xs1, us1), info = om.TargetSelectorOracle(**kw, ylb=ylb, yub=yub).solve(ysp, d, return_info=True)
        err = np.max(np.abs(u - us1.ravel())) / max(1.0, np.max(np.abs(us1)))
        worst = max(worst, err)
        n_act_tot += info.get("n_active", 0)
        assert err <= 1e-7, (trial, err, it, info)
    print(f"200 random output-constrained target problems: worst |us - oracle| / max(1,|us|) = {worst:.2e}, "
          f"{n_act_tot} active constraints in total, {n_inf} infeasible (flagged by both)")


if __name__ == "__main__":
    main()
