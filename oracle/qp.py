"""Exact / reference-style dense QP solvers in NumPy float64 (TEST INFRASTRUCTURE).

The reference hands every QP to ``cvxopt.solvers.qp`` (linearMPC.py:304-305, :503-504), a
third-party primal-dual interior-point code that is absent here (version unpinned upstream, not
installable offline).  Two stand-ins live in this file:

* exact active-set solvers (``BoxQP``, ``solve_box_qp``, ``solve_eq_box_qp``): the QPs are
  strictly convex on the feasible subspace, so the minimiser is unique and solver independent;
  these return it with KKT residual ~1e-13 and are what parity tests compare against;
* ``ipm_qp``: a dense primal-dual interior-point method with the *cost structure* of cvxopt's
  ``coneqp`` on these problems (dense ``G``, one ``P + G' W^-2 G`` product and one Cholesky per
  iteration, cold start, cvxopt default tolerances).  It is the timed CPU baseline.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def box_kkt_residual(P, q, u, lb, ub):
    """Natural-map KKT residual ||u - clip(u - (Pu+q))||_inf of a box QP."""
    g = P @ u + q
    return float(np.max(np.abs(u - np.clip(u - g, lb, ub))))


class BoxQP:
    """min 1/2 u'Pu + q'u, lb <= u <= ub for a fixed P > 0; factorises P once.

    Primal-dual active set (semismooth Newton) in Schur-complement form on P^-1, with a
    Douglas-Rachford/ADMM + polish fallback should the active set cycle.
    """

    def __init__(self, P):
        self.P = np.asarray(P, dtype=np.float64)
        self.n = self.P.shape[0]
        self.cho = scipy.linalg.cho_factor(self.P, lower=True)
        self._Pinv = None
        self._dr = None

    @property
    def Pinv(self):
        if self._Pinv is None:
            Pi = scipy.linalg.cho_solve(self.cho, np.eye(self.n))
            self._Pinv = 0.5 * (Pi + Pi.T)
        return self._Pinv

    def _polish(self, q, u_unc, lb, ub, lower, upper):
        """Equality-constrained solve with the given active sets; returns (u, mu)."""
        act = lower | upper
        u = u_unc.copy()
        mu = np.zeros(self.n)
        if act.any():
            idx = np.flatnonzero(act)
            bnd = np.where(lower, lb, ub)[idx]
            S = self.Pinv[np.ix_(idx, idx)]
            lam = scipy.linalg.solve(S, u_unc[idx] - bnd, assume_a="pos")
            u = u_unc - self.Pinv[:, idx] @ lam
            u[idx] = bnd
            mu[idx] = lam
        return u, mu

    def solve(self, q, lb, ub, tol=1e-12, max_pdas=60):
        q = np.asarray(q, dtype=np.float64).reshape(-1)
        lb = np.broadcast_to(np.asarray(lb, dtype=np.float64).reshape(-1), (self.n,))
        ub = np.broadcast_to(np.asarray(ub, dtype=np.float64).reshape(-1), (self.n,))
        u_unc = -scipy.linalg.cho_solve(self.cho, q)
        lower = u_unc < lb
        upper = u_unc > ub
        seen = set()
        info = dict(method="pdas", iters=0)
        u = np.clip(u_unc, lb, ub)
        for it in range(max_pdas):
            u, mu = self._polish(q, u_unc, lb, ub, lower, upper)
            # mu >= 0 wanted on upper, <= 0 on lower (P u + q + mu = 0)
            new_upper = (upper & (mu > 0)) | (~(lower | upper) & (u > ub))
            new_lower = (lower & (mu < 0)) | (~(lower | upper) & (u < lb))
            info["iters"] = it + 1
            if np.array_equal(new_upper, upper) and np.array_equal(new_lower, lower):
                break
            key = (new_lower.tobytes(), new_upper.tobytes())
            if key in seen:
                u = self._fallback(q, lb, ub, u_unc, tol)
                info["method"] = "dr+polish"
                break
            seen.add(key)
            lower, upper = new_lower, new_upper
        else:
            u = self._fallback(q, lb, ub, u_unc, tol)
            info["method"] = "dr+polish"
        u = np.clip(u, lb, ub)
        res = box_kkt_residual(self.P, q, u, lb, ub)
        if res > tol:
            u = self._refine(q, lb, ub, u)
            res = box_kkt_residual(self.P, q, u, lb, ub)
        info["kkt"] = res
        info["cost"] = float(0.5 * u @ (self.P @ u) + q @ u)
        info["n_active"] = int(np.sum((u <= lb) | (u >= ub)))
        return u, info

    def _refine(self, q, lb, ub, u):
        """Direct re-solve on the current active set using P itself (not P^-1)."""
        g = self.P @ u + q
        lower = (u <= lb) & (g > 0)
        upper = (u >= ub) & (g < 0)
        free = ~(lower | upper)
        v = np.where(lower, lb, np.where(upper, ub, 0.0))
        if free.any():
            f = np.flatnonzero(free)
            a = np.flatnonzero(~free)
            rhs = -(q[f] + self.P[np.ix_(f, a)] @ v[a])
            v[f] = scipy.linalg.solve(self.P[np.ix_(f, f)], rhs, assume_a="pos")
        return np.clip(v, lb, ub)

    def _fallback(self, q, lb, ub, u_unc, tol, rho=None, alpha=1.6, max_iter=20000):
        if self._dr is None:
            w = np.linalg.eigvalsh(self.P)
            rho = float(np.sqrt(w[0] * w[-1])) if rho is None else rho
            Mi = scipy.linalg.inv(self.P + rho * np.eye(self.n))
            self._dr = (rho, 0.5 * (Mi + Mi.T))
        rho, Mi = self._dr
        c = Mi @ q
        v = u_unc.copy()
        for it in range(max_iter):
            z = np.clip(v, lb, ub)
            x = rho * (Mi @ (2 * z - v)) - c
            v = v + alpha * (x - z)
            if it % 25 == 24:
                z = np.clip(v, lb, ub)
                g = self.P @ z + q
                lower = (z <= lb) & (g > 0)
                upper = (z >= ub) & (g < 0)
                u, _ = self._polish(q, u_unc, lb, ub, lower, upper)
                if np.all(u >= lb - 1e-13) and np.all(u <= ub + 1e-13):
                    u = np.clip(u, lb, ub)
                    if box_kkt_residual(self.P, q, u, lb, ub) <= tol:
                        return u
        return np.clip(v, lb, ub)


_BOX_CACHE: dict = {}


def solve_box_qp(P, q, lb, ub, tol=1e-12):
    """Exact box-QP solve; ``P`` factorisations are cached per array object."""
    key = id(P)
    ent = _BOX_CACHE.get(key)
    if ent is None or ent[0] is not P:
        if len(_BOX_CACHE) > 8:
            _BOX_CACHE.clear()
        ent = (P, BoxQP(P))
        _BOX_CACHE[key] = ent
    u, info = ent[1].solve(q, lb, ub, tol=tol)
    return u.reshape(-1, 1), info


def solve_eq_box_qp(P, q, Aeq, b, lb, ub, max_iter=100):
    """min 1/2 w'Pw + q'w  s.t. Aeq w = b, lb <= w <= ub (infinite bounds allowed).

    Primal-dual active set on the KKT system; P may be singular as long as it is positive
    definite on null(Aeq) restricted to the free variables (true for the target selector).
    """
    P = np.asarray(P, float)
    n, m = P.shape[0], Aeq.shape[0]
    q = np.asarray(q, float).reshape(-1)
    b = np.asarray(b, float).reshape(-1)
    lb = np.asarray(lb, float).reshape(-1)
    ub = np.asarray(ub, float).reshape(-1)
    lower = np.zeros(n, bool)
    upper = np.zeros(n, bool)
    seen = set()
    w = np.zeros(n)
    for it in range(max_iter):
        act = lower | upper
        f = np.flatnonzero(~act)
        a = np.flatnonzero(act)
        w = np.where(lower, lb, np.where(upper, ub, 0.0))
        nf = f.size
        K = np.zeros((nf + m, nf + m))
        K[:nf, :nf] = P[np.ix_(f, f)]
        K[:nf, nf:] = Aeq[:, f].T
        K[nf:, :nf] = Aeq[:, f]
        rhs = np.concatenate([-(q[f] + P[np.ix_(f, a)] @ w[a]), b - Aeq[:, a] @ w[a]])
        sol = scipy.linalg.solve(K, rhs, assume_a="sym")
        w[f] = sol[:nf]
        nu = sol[nf:]
        mu = -(P @ w + q + Aeq.T @ nu)          # P w + q + Aeq' nu + mu = 0
        new_upper = (upper & (mu > 0)) | (~act & (w > ub))
        new_lower = (lower & (mu < 0)) | (~act & (w < lb))
        if np.array_equal(new_upper, upper) and np.array_equal(new_lower, lower):
            info = dict(iters=it + 1, n_active=int(act.sum()),
                        cost=float(0.5 * w @ (P @ w) + q @ w),
                        eq_res=float(np.max(np.abs(Aeq @ w - b))) if m else 0.0,
                        stat_res=float(np.max(np.abs((P @ w + q + Aeq.T @ nu)[~act]))) if nf else 0.0)
            return w.reshape(-1, 1), info
        key = (new_lower.tobytes(), new_upper.tobytes())
        if key in seen:
            return _eq_box_reduced(P, q, Aeq, b, lb, ub)
        seen.add(key)
        lower, upper = new_lower, new_upper
    return _eq_box_reduced(P, q, Aeq, b, lb, ub)


def _eq_box_reduced(P, q, Aeq, b, lb, ub):
    """Fallback when the KKT active set cycles: eliminate the unbounded variables through the
    equalities (needs as many equalities as unbounded variables, square and invertible - the
    target selector with H empty) and solve the remaining box QP with ``BoxQP``."""
    n, m = P.shape[0], Aeq.shape[0]
    free = ~(np.isfinite(lb) | np.isfinite(ub))
    fi, bi = np.flatnonzero(free), np.flatnonzero(~free)
    if fi.size != m:
        raise RuntimeError("solve_eq_box_qp: active set cycled and the problem is not reducible")
    A1, A2 = Aeq[:, fi], Aeq[:, bi]
    T = -np.linalg.solve(A1, A2)                  # w_free = T w_box + t0
    t0 = np.linalg.solve(A1, b)
    Pff, Pfb, Pbb = P[np.ix_(fi, fi)], P[np.ix_(fi, bi)], P[np.ix_(bi, bi)]
    Hr = T.T @ Pff @ T + T.T @ Pfb + Pfb.T @ T + Pbb
    Hr = 0.5 * (Hr + Hr.T)
    fr = T.T @ (Pff @ t0 + q[fi]) + Pfb.T @ t0 + q[bi]
    ub_, info = BoxQP(Hr).solve(fr, lb[bi], ub[bi], tol=1e-13)
    w = np.zeros(n)
    w[bi] = ub_
    w[fi] = T @ ub_ + t0
    info = dict(info, method="reduced:" + info["method"], eq_res=float(np.max(np.abs(Aeq @ w - b))),
                cost=float(0.5 * w @ (P @ w) + q @ w))
    return w.reshape(-1, 1), info


def ipm_qp(P, q, G, h, abstol=1e-7, reltol=1e-6, feastol=1e-7, maxiters=100,
           diagonal_G=False, rematerialise=True):
    """Dense primal-dual (Mehrotra) interior point for min 1/2x'Px+q'x s.t. Gx<=h.

    Cost structure mirrors cvxopt ``coneqp`` with the default 'chol2'-style KKT solve on a
    problem with only linear inequalities: per iteration one scaled copy of dense ``G``, one
    ``P + G' D G`` product, one Cholesky and two back-solves; stopping rule = cvxopt defaults
    (no options are set anywhere in the reference).  ``rematerialise`` copies P and G on entry,
    as ``array_to_matrix`` does on every reference call (linearMPC.py:15-20, :503).
    ``diagonal_G=True`` is the structure-exploiting variant (G = [I;-I] never formed).
    Returns (x, info).
    """
    P = np.array(P, dtype=np.float64, copy=rematerialise)
    q = np.asarray(q, float).reshape(-1)
    h = np.asarray(h, float).reshape(-1)
    n = P.shape[0]
    if diagonal_G:
        m = 2 * n
        Gmul = lambda x: np.concatenate([x, -x])
        GTmul = lambda z: z[:n] - z[n:]
    else:
        G = np.array(G, dtype=np.float64, copy=rematerialise)
        m = G.shape[0]
        Gmul = lambda x: G @ x
        GTmul = lambda z: G.T @ z

    def kkt_factor(d):
        if diagonal_G:
            H = P.copy()
            H[np.diag_indices(n)] += d[:n] + d[n:]
        else:
            H = P + G.T @ (d[:, None] * G)
        return scipy.linalg.cho_factor(H, lower=True, overwrite_a=True, check_finite=False)

    # cvxopt-style initial point: solve with identity scaling, then shift s,z into the cone
    cf = kkt_factor(np.ones(m))
    x = scipy.linalg.cho_solve(cf, -q + GTmul(h), check_finite=False)
    z = Gmul(x) - h
    s = -z
    ts = np.max(-s)
    s = s + (1.0 + ts) if ts >= 0 else s
    tz = np.max(-z)
    z = z + (1.0 + tz) if tz >= 0 else z
    resx0 = max(1.0, np.linalg.norm(q))
    resz0 = max(1.0, np.linalg.norm(h))
    info = dict(status="maxiters", iters=maxiters)
    for it in range(maxiters):
        rx = P @ x + q + GTmul(z)
        rz = Gmul(x) + s - h
        gap = float(s @ z)
        pcost = float(0.5 * x @ (P @ x) + q @ x)
        dcost = pcost + float(z @ (Gmul(x) - h))
        if pcost < 0:
            relgap = gap / -pcost
        elif dcost > 0:
            relgap = gap / dcost
        else:
            relgap = np.inf
        pres = np.linalg.norm(rz) / resz0
        dres = np.linalg.norm(rx) / resx0
        if pres <= feastol and dres <= feastol and (gap <= abstol or relgap <= reltol):
            info = dict(status="optimal", iters=it)
            break
        d = z / s
        try:
            cf = kkt_factor(d)
        except np.linalg.LinAlgError:          # z/s has left the range a Cholesky survives: as converged as FP64 allows
            info = dict(status="numerical", iters=it)
            break
        mu = gap / m

        def step(rc):
            # (P + G'DG) dx = -rx - G'( (rc - z*rz)/s ),  dz = (rc - z*rz... )
            t = (z * rz - rc) / s
            dx = scipy.linalg.cho_solve(cf, -rx - GTmul(t), check_finite=False)
            dz = t + d * Gmul(dx)
            ds = -(rc + s * dz) / z
            return dx, ds, dz

        dx, ds, dz = step(s * z)                        # affine scaling
        amax = lambda v, dv: min(1.0, 1.0 / max(1e-300, np.max(-dv / v)))
        a = min(amax(s, ds), amax(z, dz))
        sigma = (1.0 - a) ** 3
        dx, ds, dz = step(s * z + ds * dz - sigma * mu)  # combined direction
        a = min(1.0, 0.99 / max(1e-300, max(np.max(-ds / s), np.max(-dz / z))))
        x = x + a * dx
        s = s + a * ds
        z = z + a * dz
    info["gap"] = float(s @ z)
    return x.reshape(-1, 1), info


def solve_general_qp(P, q, G, h):
    """Inequality QP (re-parameterised regulator, linearMPC.py:476-493): interior point to a tight tolerance,
    then an active-set polish - the constraints the interior point identifies as active are imposed as equalities
    and the KKT system is solved directly; the polished point is kept when it is feasible with multipliers of the
    right sign (then it is the exact minimiser), otherwise the interior-point answer stands."""
    P, G = np.asarray(P, float), np.asarray(G, float)
    q, h = np.asarray(q, float).reshape(-1), np.asarray(h, float).reshape(-1)
    x, info = ipm_qp(P, q, G, h, abstol=1e-13, reltol=1e-13, feastol=1e-12, maxiters=200, rematerialise=False)
    x = x.reshape(-1)
    slack = h - G @ x
    act = np.flatnonzero(slack <= 1e-7 * (1.0 + np.abs(h)))
    if act.size:
        Ga = G[act]
        # drop linearly dependent rows (e.g. u <= ub and -u <= -lb with lb == ub)
        _, r, piv = scipy.linalg.qr(Ga.T, mode="economic", pivoting=True)
        rank = int(np.sum(np.abs(np.diag(r)) > 1e-10 * max(1.0, abs(r[0, 0]))))
        act = act[np.sort(piv[:rank])]
        Ga = G[act]
        n, k = P.shape[0], act.size
        K = np.block([[P, Ga.T], [Ga, np.zeros((k, k))]])
        try:
            sol = scipy.linalg.solve(K, np.concatenate([-q, h[act]]), assume_a="sym")
            xp, lam = sol[:n], sol[n:]
            if np.all(G @ xp <= h + 1e-10 * (1.0 + np.abs(h))) and np.all(lam >= -1e-9 * (1.0 + np.abs(lam).max())):
                x = xp
                info = dict(info, polished=True, n_active=int(k))
        except np.linalg.LinAlgError:
            pass
    info["cost"] = float(0.5 * x @ (P @ x) + q @ x)
    return x.reshape(-1, 1), info
