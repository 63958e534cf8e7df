"""GPU: the output-constrained target selector (k_ts_general, dual active set, one warp per sample) against the oracle's
(xs, us)-space solve of the reference formulation (lib/linearMPC.py:229-311 with ylb / yub), alone and inside the
closed-loop engine."""
import numpy as np
import pytest

from oracle import linear_mpc as om

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def _pair(p, ylb, yub):
    from industrial_nnmpc_2021_b200.linearMPC import TargetSelector
    kw = dict(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs, ulb=p.ulb, uub=p.uub, ylb=ylb, yub=yub)
    return TargetSelector(**kw), om.TargetSelectorOracle(**kw)


@pytest.mark.parametrize("which", ["cstrs", "cdu_small"])
def test_output_constrained_targets_match_oracle(torch_cuda, which, cstrs_problem, cdu_small_problem):
    p = cstrs_problem if which == "cstrs" else cdu_small_problem
    free = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs,
                                   ulb=p.ulb, uub=p.uub)
    idx = np.arange(0, 60000, 2500)
    YSP, D = p.setpoints[idx], p.disturbances[idx]
    # output box around the input-constrained optima (every sample's own optimum lies inside), then two outputs are cut
    # at the median of their optima: about half of the samples have their optimum cut off (Ny > Nu for the CDU family,
    # so a tight box in every output would be empty)
    Y0 = np.array([(p.C @ free.solve(y[:, None], d[:, None])[0] + p.Cd @ d[:, None])[:, 0] for y, d in zip(YSP, D)])
    span = Y0.max(axis=0) - Y0.min(axis=0)
    ylb, yub = (Y0.min(axis=0) - 0.05 * span - 1e-3)[:, None], (Y0.max(axis=0) + 0.05 * span + 1e-3)[:, None]
    j1, j2 = np.argsort(span)[-1], np.argsort(span)[-2]
    if p.Ny <= p.Nu:
        yub[j1] = np.quantile(Y0[:, j1], 0.6)
        ylb[j2] = np.median(Y0[:, j2])
    else:
        # CDU family: with Rs = 1e-6, Qs = 1e-16 the targets sit in corners of the input box and most outputs cannot be
        # moved without leaving it (a cut there is simply infeasible); output 3 can
        yub[3] = np.quantile(Y0[:, 3], 0.6)
    ts, ots = _pair(p, ylb, yub)
    assert ts.h is None and np.allclose(ts.G, ots.G)
    keep, ref = [], []
    for i in range(YSP.shape[0]):
        try:
            ref.append(ots.solve(YSP[i][:, None], D[i][:, None], return_info=True))
            keep.append(i)
        except ValueError:
            pass                                  # infeasible for this (ysp, d): covered below
    assert len(keep) >= 4
    xs, us, it = ts.solve_batch(YSP[keep], D[keep], return_iters=True)
    n_out = 0
    for k, ((oxs, ous), info) in enumerate(ref):
        y = p.C @ xs[k] + p.Cd @ D[keep[k]]
        assert np.all(y >= ylb[:, 0] - 1e-8) and np.all(y <= yub[:, 0] + 1e-8)
        assert np.all(us[k] >= p.ulb[:, 0] - 1e-12) and np.all(us[k] <= p.uub[:, 0] + 1e-12)
        w = np.concatenate([xs[k], us[k]])
        b = ots.tb @ np.concatenate([YSP[keep[k]], D[keep[k]]])
        assert np.max(np.abs(ots.tA @ w - b)) <= 1e-8
        q = ots.changing(YSP[keep[k]][:, None], D[keep[k]][:, None])[0][:, 0]
        ow = np.vstack([oxs, ous])[:, 0]
        c, oc = 0.5 * w @ ots.P @ w + q @ w, 0.5 * ow @ ots.P @ ow + q @ ow
        assert abs(c - oc) <= 1e-7 * max(1.0, abs(oc)), (k, c, oc)
        assert np.max(np.abs(us[k] - ous[:, 0])) <= 1e-6 * max(1.0, np.abs(ous).max()), k
        assert np.max(np.abs(xs[k] - oxs[:, 0])) <= 1e-5 * max(1.0, np.abs(oxs).max()), k
        n_out += int(np.any(np.abs(y - ylb[:, 0]) <= 1e-7) or np.any(np.abs(y - yub[:, 0]) <= 1e-7))
    assert n_out >= 2 and it.min() >= 0
    # single-solve signature and history, as the reference class
    oxs1, ous1 = ts.solve(YSP[keep[0]][:, None], D[keep[0]][:, None])
    assert oxs1.shape == (p.Nx, 1) and ous1.shape == (p.Nu, 1) and len(ts.us) == 1


def test_infeasible_target_raises(torch_cuda, cdu_small_problem):
    from industrial_nnmpc_2021_b200._lib import NnmpcError
    p = cdu_small_problem
    ts, ots = _pair(p, np.full((p.Ny, 1), 50.0), np.full((p.Ny, 1), 51.0))      # no input in the box reaches these outputs
    with pytest.raises(ValueError):
        ots.solve(p.setpoints[0][:, None], p.disturbances[0][:, None])
    with pytest.raises(NnmpcError):
        ts.solve_batch(p.setpoints[:3], p.disturbances[:3])


def test_closed_loop_with_output_constrained_targets(torch_cuda, cdu_small_problem):
    """The engine with an output-constrained target selector == the oracle's simulate_offline with the same bounds."""
    from industrial_nnmpc_2021_b200.linearMPC import ClosedLoopEngine, LinearMPCController
    p = cdu_small_problem
    T, starts = 25, (388, 787, 1187)            # windows around set-point changes of the PRBS scenario
    free = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs,
                                   ulb=p.ulb, uub=p.uub)
    sp = np.stack([p.setpoints[s:s + T] for s in starts])
    ds = np.stack([p.disturbances[s:s + T] for s in starts])
    Y0 = np.array([(p.C @ free.solve(sp[c, t][:, None], ds[c, t][:, None])[0] + p.Cd @ ds[c, t][:, None])[:, 0]
                   for c in range(len(starts)) for t in range(T)])
    # With the CDU tuning (Rs = 1e-6, Qs = 1e-16) the input-constrained targets sit in corners of the input box, and
    # most outputs cannot be moved at all without leaving it; output 2 can: cap it 0.2 below its largest target value
    j = 2
    ylb, yub = np.full((p.Ny, 1), -1e3), np.full((p.Ny, 1), 1e3)
    yub[j] = Y0[:, j].max() - 0.2
    ts, ots = _pair(p, ylb, yub)
    reg = LinearMPCController.setup_regulator(p.A, p.B, p.Q, p.R, p.S, p.N, p.ulb, p.uub)
    eng = ClosedLoopEngine(reg, ts, p.A, p.B, p.Bd)
    res = eng.run(p.xprior, p.uprev, sp, ds)
    assert not res["maxiter_hit"] and float(res["kkt"].max()) <= 1e-8
    oreg = om.setup_regulator(p.A, p.B, p.Q, p.R, p.S, p.N, p.ulb, p.uub)
    cut = 0
    for c in range(len(starts)):
        od = om.simulate_offline(x0=p.xprior, uprev0=p.uprev, A=p.A, B=p.B, Bd=p.Bd, regulator=oreg, ulb=p.ulb, uub=p.uub,
                                 target_selector=ots, setpoints=sp[c], disturbances=ds[c])
        for k in ("xs", "us", "u", "x"):
            err = np.max(np.abs(res[k][c] - od[k])) / max(1.0, np.max(np.abs(od[k])))
            assert err <= 1e-6, (c, k, err)
        ys = od["xs"] @ p.C.T + ds[c] @ p.Cd.T
        cut += int(np.sum(np.abs(ys[:, j] - yub[j, 0]) <= 1e-7))
    assert cut >= 3
