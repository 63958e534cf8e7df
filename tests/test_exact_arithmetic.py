"""CPU checks (NumPy) of the arithmetic identities the mixed-precision engine relies on:
  * the error-free base-128 slicing behind the INT8 tensor-core exact GEMM (csrc/oz_gemm.cuh): digits are exact,
    integer accumulation is exact, the only error is the level truncation and it obeys the documented bound;
  * the re-anchoring identity of a failed KKT check (csrc/lp_iter.cuh k_reanchor): w_lp = z + g/rho gives
    Top w_lp - c = z exactly.
They restate the device code line by line; the GPU tests compare the kernels themselves against FP64 NumPy."""
import numpy as np
import pytest


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


# ---- the fp16-increment Douglas-Rachford pass (csrc/lp_iter.cuh), restated in NumPy -----------------------------
def _pow2_scale(est):
    """lp.cuh pow2_scale: largest power of two s with s * est in [8, 16], clamped to 2^+-60."""
    if not est > 0.0:
        return 2.0 ** 60
    ex = np.frexp(est)[1]
    return float(np.ldexp(1.0, int(np.clip(4 - ex, -60, 60))))


def _quantise(dw, s):
    """quantise_dw: fp16 increment with a power-of-two row scale; the residual (fp32) carries what did not fit."""
    t = np.clip(dw * s, -60000.0, 60000.0)
    q = t.astype(np.float16)
    e = (dw - q.astype(np.float64) / s).astype(np.float32)
    return q, e


def _mixed_solve(P, q, lb, ub, passes, reanchor_at=(), t2_every=0):
    """One QP through the mixed-precision tiers: exact anchor, fp16 increments against a two-term fp16 operator
    split with fp32 accumulation, FP64 state; optional re-anchoring from the exact gradient (k_reanchor/k_lp_emit).
    t2_every = m > 0: the deferred second term of lp_iter.cuh - a pass multiplies T1 only, the pending sum S of the
    increments (fp16, own power-of-two scale, started at 1/8 of the operand scale) is delivered through T2 every m-th
    pass (EpiAddX) - on pass 0 right after the anchor - and restarted."""
    n = len(q)
    lam = np.linalg.eigvalsh(P)
    rho = 0.5 * np.sqrt(lam[0] * lam[-1]) * np.diag(P) / np.exp(np.mean(np.log(np.diag(P))))
    Minv = np.linalg.inv(P + np.diag(rho))
    Top, c, alpha = Minv * rho[None, :], Minv @ q, 1.8
    sT = float(np.ldexp(1.0, 10 - np.frexp(np.abs(Top).max())[1]))
    T1 = (sT * Top).astype(np.float16)
    T2 = (sT * Top - T1.astype(np.float64)).astype(np.float16)
    Tlp = T1.astype(np.float32) + T2.astype(np.float32)          # both products land in one fp32 accumulator
    T1f, T2f = T1.astype(np.float32), T2.astype(np.float32)
    clip = lambda v: np.minimum(np.maximum(v, lb), ub)
    v = -np.linalg.solve(P, q)                                    # cold start: the unconstrained law
    # exact anchor + one full-precision step + first increment (k_anchor_prep, anchor GEMM, k_dr_first)
    w_lp = 2 * clip(v) - v
    x = Top @ w_lp - c
    d = x - clip(v)
    v = v + alpha * d
    dw = (2 * clip(v) - v) - w_lp
    s_in = _pow2_scale(np.abs(dw).max())
    dq, e = _quantise(dw, s_in)
    s_out = _pow2_scale(3 * alpha * np.abs(d).max())
    S, sS = (0.125 * dq.astype(np.float32)).astype(np.float16), 0.125 * s_in      # k_dr_first: the pending sum starts
    hist = []
    for k in range(passes):
        if k in reanchor_at:                                     # a failed check: x := z exactly for w_lp = z + g / rho
            z = clip(v)
            g = P @ z + q
            x = z.copy()
            dw = (2 * z - v) - (z + g / rho)
            s_in = _pow2_scale(np.abs(dw).max())
            dq, e = _quantise(dw, s_in)
            S, sS = (0.125 * dq.astype(np.float32)).astype(np.float16), 0.125 * s_in          # k_lp_emit
        if t2_every:
            deliver = k % t2_every == 0
            if deliver:                                          # EpiAddX: x += T2 S / (s_T s_S), S includes this pass's operand
                x = x + (T2f @ S.astype(np.float32)).astype(np.float64) / (sT * sS)
            acc = T1f @ dq.astype(np.float32)
        else:
            acc = Tlp @ dq.astype(np.float32)                    # tcgen05: fp16 x fp16 -> fp32
        x = x + acc.astype(np.float64) / (sT * s_in)
        wl = (2 * clip(v) - v) - e.astype(np.float64)
        d = x - clip(v)
        v = v + alpha * d
        dw = (2 * clip(v) - v) - wl
        dq, e = _quantise(dw, s_out)
        if t2_every:                                             # EpiDelta: S+ = [new sum] g dq  or  S + g dq, saturating
            if deliver:
                sS = 0.125 * s_out
                S = (0.125 * dq.astype(np.float32)).astype(np.float16)
            else:
                gsc = np.float32(sS / s_out)
                S = np.clip(S.astype(np.float32) + gsc * dq.astype(np.float32), -65504.0, 65504.0).astype(np.float16)
        s_in, s_out = s_out, _pow2_scale(3 * alpha * np.abs(d).max())
        z = clip(v)
        hist.append((np.abs(d).max(), np.abs(z - clip(z - (P @ z + q))).max()))
    return clip(v), hist


def test_fp16_increment_iteration_reaches_the_fp64_optimum():
    """The tensor-core pass changes HOW x = Top w - c is tracked, not the fixed point: the iterate gets far below
    what fp16 (1e-3) could resolve directly, because the increments shrink with the iteration and their
    quantisation residual is fed forward; a re-anchor from the exact gradient removes the drift that is left."""
    rng = np.random.default_rng(9)
    n = 60
    R = rng.standard_normal((n, n))
    P = R @ R.T / n + 0.3 * np.eye(n)
    q = 2.0 * rng.standard_normal(n)
    lb, ub = -0.5 * np.ones(n), 0.5 * np.ones(n)
    from oracle import qp as oq
    ue, info = oq.solve_box_qp(P, q[:, None], lb[:, None], ub[:, None])
    assert info["n_active"] > 3
    z, hist = _mixed_solve(P, q, lb, ub, passes=140)
    # a cold start makes large first increments: what the fp16 operator split and the fp32 accumulation lost on them
    # stays in x, so the iteration settles (||d|| ~ 1e-16: the cheap residual says "converged") a drift of ~1e-7
    # away from the optimum - three to four orders below fp16 resolution, but above the tolerance.  This is why every
    # returned point is checked exactly, and what a failed check repairs.
    assert hist[-1][0] <= 1e-12 and 1e-9 < hist[-1][1] <= 1e-6
    assert np.max(np.abs(z - ue[:, 0])) <= 1e-6
    # the failed check re-anchors from its own gradient (x := z exactly for w_lp = z + g / rho): what is left to
    # deliver is tiny, so is its error, and the exact KKT residual falls to rounding level
    z2, hist2 = _mixed_solve(P, q, lb, ub, passes=180, reanchor_at=(120,))
    assert hist2[-1][1] <= 1e-11, hist2[-1]
    assert np.max(np.abs(z2 - ue[:, 0])) <= 1e-10
    assert np.all(z2 >= lb) and np.all(z2 <= ub)


@pytest.mark.parametrize("m", [1, 4, 8])
def test_deferred_second_operator_term_keeps_fixed_points_and_pace(m):
    """One fp16 operator term per pass, the second delivered every m-th pass from the pending sums: the iteration has
    the fixed points of the two-term form (nothing is lost, only delayed - the drift left at ||d|| ~ 0 is the same
    ~1e-7 the exact check repairs), converges at the same pace, and re-anchors to rounding level like it."""
    rng = np.random.default_rng(9)
    n = 60
    R = rng.standard_normal((n, n))
    P = R @ R.T / n + 0.3 * np.eye(n)
    q = 2.0 * rng.standard_normal(n)
    lb, ub = -0.5 * np.ones(n), 0.5 * np.ones(n)
    from oracle import qp as oq
    ue, _ = oq.solve_box_qp(P, q[:, None], lb[:, None], ub[:, None])
    _, h2 = _mixed_solve(P, q, lb, ub, passes=140)
    z1, h1 = _mixed_solve(P, q, lb, ub, passes=140, t2_every=m)
    first = lambda h, thr: next(k for k, (d, _) in enumerate(h) if d <= thr)
    for thr in (1e-4, 1e-7, 1e-10):
        assert first(h1, thr) <= first(h2, thr) + 3, (m, thr, first(h1, thr), first(h2, thr))
    assert h1[-1][0] <= 1e-12 and h1[-1][1] <= 1e-6 and np.max(np.abs(z1 - ue[:, 0])) <= 1e-6
    z3, h3 = _mixed_solve(P, q, lb, ub, passes=180, reanchor_at=(120,), t2_every=m)
    # (the fp16 rounding of the pending sums adds to the rounding-level residue: 2e-11 here, tolerance 1e-9)
    assert h3[-1][1] <= 1e-10 and np.max(np.abs(z3 - ue[:, 0])) <= 1e-9
    # skipping the second term WITHOUT delivering it later is not the same thing: the drift grows by orders
    _, hdrop = _mixed_solve(P, q, lb, ub, passes=140, t2_every=10 ** 9)      # delivery on pass 0 only
    assert hdrop[-1][1] > 20 * h1[-1][1]
