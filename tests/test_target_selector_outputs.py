"""Output-constrained target selector (lib/linearMPC.py:242-248, :284-288): the oracle and the product's host
formulation against a fixture generated from the reference's own TargetSelector, the oracle's solve against an
independent feasibility verdict, and the NumPy restatement of the GPU kernel's dual active-set method against the
oracle (CPU only; the kernel itself is tested in test_gpu_target_selector_outputs.py)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import linear_mpc as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_ts():
    with np.load(os.path.join(ROOT, "tests", "golden", "target_selector_outputs.npz")) as z:
        return {k: z[k] for k in z.files}


def _oracle(p, ylb, yub):
    return om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs, ulb=p.ulb,
                                   uub=p.uub, ylb=ylb, yub=yub)


def test_oracle_formulation_matches_reference(golden_ts, cdu_small_problem):
    g, p = golden_ts, cdu_small_problem
    ts = _oracle(p, g["ylb"], g["yub"])
    assert ts.h is None
    for k in ("P", "G", "f", "e", "F", "tA", "tb"):
        assert np.allclose(getattr(ts, k), g[k], atol=1e-14), k
    q, h, b = ts.changing(g["ysp"], g["d"])
    for k, v in (("q", q), ("h", h), ("b", b)):
        assert v.shape == g[k].shape and np.allclose(v, g[k], atol=1e-13), k


def test_host_formulation_matches_reference(golden_ts, cdu_small_problem):
    """The product class's reference attributes for this branch (no GPU: the handle is never created)."""
    from industrial_nnmpc_2021_b200.linearMPC import TargetSelector
    g, p = golden_ts, cdu_small_problem
    ts = TargetSelector.__new__(TargetSelector)
    ts.A, ts.B, ts.C, ts.H, ts.Bd, ts.Cd, ts.Rs, ts.Qs = p.A, p.B, p.C, p.H, p.Bd, p.Cd, p.Rs, p.Qs
    ts.Nx, ts.Nu = p.B.shape
    ts.Ny, ts.Nd, ts.Nz = p.C.shape[0], p.Bd.shape[1], p.H.shape[0]
    ts.usp, ts.ulb, ts.uub, ts.ylb, ts.yub = p.usp, p.ulb, p.uub, g["ylb"], g["yub"]
    ts._setup_fixed_matrices()
    assert ts.h is None
    for k in ("P", "G", "f", "e", "F", "tA", "tb"):
        assert np.allclose(getattr(ts, k), g[k], atol=1e-14), k
    q, h, b = ts._setup_changing_matrices(g["ysp"], g["d"])
    assert np.allclose(q, g["q"], atol=1e-13) and np.allclose(h, g["h"], atol=1e-13) and np.allclose(b, g["b"], atol=1e-13)
    # reduced operators of the GPU solve: Mbar = Abar Ht^-1 Abar', rows [C Gx; I]
    assert ts.Abar.shape == (ts.Ny + ts.Nu, ts.Nu) and np.allclose(ts.Hinv @ ts.Ht, np.eye(ts.Nu), atol=1e-8)
    assert np.allclose(ts.Mbar, ts.Abar @ np.linalg.solve(ts.Ht, ts.Abar.T), rtol=1e-9, atol=1e-9 * np.abs(ts.Mbar).max())
    # outputs of the reduced problem are the outputs of the reference's: C xs + Cd d = CG us + Ryd d
    rng = np.random.default_rng(0)
    us, d = rng.standard_normal((ts.Nu, 1)), rng.standard_normal((ts.Nd, 1))
    xs = ts.Gx @ us + ts.Gd @ d
    assert np.allclose(p.C @ xs + p.Cd @ d, ts.Abar[:ts.Ny] @ us + ts.Ryd @ d, atol=1e-12)


def test_oracle_solution_is_a_kkt_point(golden_ts, cdu_small_problem):
    g, p = golden_ts, cdu_small_problem
    ts = _oracle(p, g["ylb"], g["yub"])
    free = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs,
                                   ulb=p.ulb, uub=p.uub)
    rng = np.random.default_rng(5)
    n_tight = 0
    for _ in range(6):
        ysp, d = 0.8 * rng.standard_normal((p.Ny, 1)), 0.3 * rng.standard_normal((p.Nd, 1))
        xs0, us0 = free.solve(ysp, d)
        y0 = p.C @ xs0 + p.Cd @ d
        # bounds around the input-constrained optimum, one of them cutting it off
        ylb, yub = y0 - 0.5, y0 + 0.5
        yub[2] = y0[2] - 0.02
        t = _oracle(p, ylb, yub)
        (xs, us), info = t.solve(ysp, d, return_info=True)
        q, h, b = t.changing(ysp, d)
        w = np.vstack([xs, us])
        assert np.max(np.abs(t.tA @ w - b)) <= 1e-9 and np.max(t.G @ w - h) <= 1e-9
        assert abs((p.C @ xs + p.Cd @ d)[2, 0] - yub[2, 0]) <= 1e-8        # the cut is active
        c0 = (0.5 * w.T @ t.P @ w + q.T @ w).item()
        w0 = np.vstack([xs0, us0])
        assert c0 >= (0.5 * w0.T @ t.P @ w0 + q.T @ w0).item() - 1e-12                 # a constrained optimum costs more
        n_tight += info.get("n_active", 0)
    assert n_tight >= 6
    with pytest.raises(ValueError):                                                # empty output box
        _oracle(p, np.full((p.Ny, 1), 50.0), np.full((p.Ny, 1), 51.0)).solve(g["ysp"], g["d"])
    assert ts.solve(g["ysp"], g["d"])[1].shape == (p.Nu, 1)


def test_dual_active_set_restatement_matches_oracle():
    """tools/probes/ts_dual_active_set.py restates k_ts_general step by step in NumPy; its main() checks 200 random
    problems against the oracle and an LP feasibility verdict."""
    spec = importlib.util.spec_from_file_location("ts_das", os.path.join(ROOT, "tools", "probes", "ts_dual_active_set.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
