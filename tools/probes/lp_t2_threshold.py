"""CPU exploration (NumPy restatement of csrc/lp_iter.cuh, see tests/test_exact_arithmetic.py): how many tensor-core
passes of a warm-started QP need the SECOND fp16 operator term T2?  Once ||d|| is below a threshold the remaining
travel of the operand is so small that the 11-bit operator T1 alone keeps x accurate; the passes after that could
run with half the MMA work and half the operand traffic (rows sorted by phase so whole tiles skip T2).

    python tools/probes/lp_t2_threshold.py

Result on the 60-variable box QP used below (round 1): with the threshold at 1e-7 the final exact KKT residual is
unchanged (1.6e-10) and T2 is used in 20 of 48 passes (warm-start perturbation 1e-3) / 30 of 58 (1e-2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_exact_arithmetic as t   # noqa: E402


def mixed(P, q, lb, ub, v0, passes, thr):
    lam = np.linalg.eigvalsh(P)
    rho = 0.5 * np.sqrt(lam[0] * lam[-1]) * np.diag(P) / np.exp(np.mean(np.log(np.diag(P))))
    Minv = np.linalg.inv(P + np.diag(rho))
    Top, c, alpha = Minv * rho[None, :], Minv @ q, 1.8
    sT = float(np.ldexp(1.0, 10 - np.frexp(np.abs(Top).max())[1]))
    T1 = (sT * Top).astype(np.float16)
    T2 = (sT * Top - T1.astype(np.float64)).astype(np.float16)
    T1f, T2f = T1.astype(np.float32), T2.astype(np.float32)
    clip = lambda v: np.minimum(np.maximum(v, lb), ub)
    v = v0.copy()
    w_lp = 2 * clip(v) - v
    x = Top @ w_lp - c
    d = x - clip(v)
    v = v + alpha * d
    dw = (2 * clip(v) - v) - w_lp
    s_in = t._pow2_scale(np.abs(dw).max())
    dq, e = t._quantise(dw, s_in)
    s_out = t._pow2_scale(3 * alpha * np.abs(d).max())
    dmax, hist, used_t2 = np.abs(d).max(), [], 0
    for _ in range(passes):
        use2 = dmax >= thr
        used_t2 += use2
        acc = T1f @ dq.astype(np.float32)
        if use2:
            acc = acc + T2f @ dq.astype(np.float32)
        x = x + acc.astype(np.float64) / (sT * s_in)
        wl = (2 * clip(v) - v) - e.astype(np.float64)
        d = x - clip(v)
        v = v + alpha * d
        dw = (2 * clip(v) - v) - wl
        dq, e = t._quantise(dw, s_out)
        s_in, s_out = s_out, t._pow2_scale(3 * alpha * np.abs(d).max())
        dmax = np.abs(d).max()
        z = clip(v)
        hist.append((dmax, np.abs(z - clip(z - (P @ z + q))).max()))
    return clip(v), v, hist, used_t2


def main():
    rng = np.random.default_rng(9)
    n = 60
    R = rng.standard_normal((n, n))
    P = R @ R.T / n + 0.3 * np.eye(n)
    lb, ub = -0.5 * np.ones(n), 0.5 * np.ones(n)
    q = 2.0 * rng.standard_normal(n)
    _, v, _, _ = mixed(P, q, lb, ub, -np.linalg.solve(P, q), 200, 0.0)
    for pert in (1e-2, 1e-3):
        q2 = q + pert * rng.standard_normal(n)
        for thr in (0.0, 1e-9, 1e-7, 1e-6, 1e-5, 1e30):
            _, _, h, used = mixed(P, q2, lb, ub, v, 80, thr)
            conv = next((k for k, (d, _) in enumerate(h) if d <= 1e-11), None)
            print(f"perturbation {pert:g}  T2 while ||d|| >= {thr:g}: T2 passes {used:3d}, ||d|| <= 1e-11 at pass {conv}, "
                  f"final exact KKT {h[-1][1]:.2e}")


if __name__ == "__main__":
    main()
