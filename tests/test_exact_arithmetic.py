"""CPU checks (NumPy) of the arithmetic identities the mixed-precision engine relies on:
  * the error-free base-128 slicing behind the INT8 tensor-core exact GEMM (csrc/oz_gemm.cuh): digits are exact,
    integer accumulation is exact, the only error is the level truncation and it obeys the documented bound;
  * the re-anchoring identity of a failed KKT check (csrc/lp_iter.cuh k_reanchor): w_lp = z + g/rho gives
    Top w_lp - c = z exactly.
They restate the device code line by line; the GPU tests compare the kernels themselves against FP64 NumPy."""
import numpy as np


def _slice_rows(A, ns):
    """k_oz_slice: per row 2^f with max|a| <= 2^(f-1), then ns signed base-128 digits in [-64, 64]."""
    m = np.abs(A).max(axis=1)
    ex = np.where(m > 0, np.frexp(m)[1], 0)
    scale = np.ldexp(1.0, ex + 1)
    t = A / scale[:, None]
    digits = []
    for _ in range(ns):
        t = t * 128.0
        d = np.rint(t)
        t = t - d
        digits.append(d.astype(np.int64))
    return digits, scale, t


def _oz_gemm(A, Bt, lmax):
    da, fa, _ = _slice_rows(A, lmax + 1)
    db, fb, _ = _slice_rows(Bt, lmax + 1)
    v = np.zeros((A.shape[0], Bt.shape[0]))
    for L in range(lmax, -1, -1):                      # Horner in 1/128, smallest level first (the epilogue's order)
        acc = np.zeros((A.shape[0], Bt.shape[0]), dtype=np.int64)
        for i in range(L + 1):
            acc += da[i] @ db[L - i].T
        assert np.abs(acc).max() < 2 ** 31             # INT32 accumulators never overflow
        v = (v + acc.astype(np.float64)) * 0.0078125
    return v * 0.0078125 * fa[:, None] * fb[None, :]


def test_slicing_is_error_free():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((7, 300)) * np.exp(rng.uniform(-20, 20, (7, 1)))
    A[3] = 0.0
    digits, scale, rem = _slice_rows(A, 8)
    assert all(np.abs(d).max() <= 64 for d in digits)
    recon = sum(d * 128.0 ** -(i + 1) for i, d in enumerate(digits)) * scale[:, None]
    # 8 digits carry 56 bits: what is left is below one unit of the last digit
    assert np.all(np.abs(rem) <= 0.5)
    assert np.all(np.abs(recon - A) <= 128.0 ** -8 * scale[:, None])
    assert np.all(recon[3] == 0.0)


def test_sliced_gemm_matches_fp64_within_the_truncation_bound():
    rng = np.random.default_rng(1)
    M, N, K = 9, 11, 700
    A = rng.standard_normal((M, K)) * np.exp(rng.uniform(-6, 2, (M, 1)))
    Bt = rng.standard_normal((N, K)) * np.exp(rng.uniform(-8, 0, (N, K)))
    ref = np.array([[float(np.dot(np.longdouble(a), np.longdouble(b))) for b in Bt] for a in A])
    scale = np.abs(A).max(axis=1, keepdims=True) * np.abs(Bt).max(axis=1)[None, :]
    for lmax, pairs in ((7, 8), (6, 7), (5, 6)):
        err = np.abs(_oz_gemm(A, Bt, lmax) - ref)
        bound = pairs * K * 2.0 ** (-7 * (lmax + 1) - 2) * 16 * scale * 1.02 + 8 * K * 2.0 ** -53 * scale
        assert np.all(err <= bound), (lmax, float((err / bound).max()))
    # 8 levels are at the level of FP64 rounding itself
    assert np.all(np.abs(_oz_gemm(A, Bt, 7) - ref) <= 4 * K * 2.0 ** -53 * (np.abs(A) @ np.abs(Bt).T) + 1e-300)


def test_failed_check_reanchors_exactly():
    rng = np.random.default_rng(2)
    n = 40
    R = rng.standard_normal((n, n))
    P = R @ R.T + 0.1 * np.eye(n)
    rho = np.exp(rng.uniform(-1, 1, n))
    Minv = np.linalg.inv(P + np.diag(rho))
    Top = Minv * rho[None, :]
    q = rng.standard_normal(n)
    c = Minv @ q
    lb, ub = -np.ones(n), np.ones(n)
    v = 2.0 * rng.standard_normal(n)
    z = np.clip(v, lb, ub)
    g = P @ z + q                                      # what the exact check computes
    w_lp = z + g / rho                                 # k_reanchor
    x = Top @ w_lp - c
    assert np.max(np.abs(x - z)) <= 1e-12 * max(1.0, np.abs(z).max()) * np.linalg.cond(P + np.diag(rho))
    # what is left to deliver vanishes at the solution: at a KKT point z = clip(z - g) and v = z - g / rho
    # is a Douglas-Rachford fixed point, so w - w_lp = (z - v) - g / rho = 0
    v_star = z - g / rho
    assert np.max(np.abs(((2 * z - v_star) - w_lp))) <= 1e-12
