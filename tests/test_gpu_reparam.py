"""GPU parity of the re-parameterised regulator path (open-loop UNSTABLE A; lib/linearMPC.py:366-382,
:476-493, :507-509): the product solves the equivalent input-space box QP on the GPU, the oracle
follows the reference literally (general-G QP in v, mapped back through tK)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import linear_mpc as om
from oracle import qp as oq

U0_RTOL = 1e-6
COST_RTOL = 1e-6
KKT_TOL = 1e-8


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def unstable_plant(nx=6, nu=2, seed=5, rho=1.04):
    """Random plant with one unstable real mode (rho) and one marginally damped pair; controllable."""
    rng = np.random.default_rng(seed)
    lam = np.concatenate([[rho, 0.97], rng.uniform(0.3, 0.9, nx - 2)])
    V = rng.standard_normal((nx, nx)) + 2.0 * np.eye(nx)
    A = V @ np.diag(lam) @ np.linalg.inv(V)
    B = rng.standard_normal((nx, nu))
    return A, B


def test_golden_unstable_case_on_gpu(torch_cuda, golden_formulation):
    """The reference-generated fixture (2 states, unstable pole 1.05): attributes in the reference's v-space
    meaning, solve() returns u.  N = 6 here (the GPU kernels want an even number of variables); the N = 5 host
    operators are checked against the fixture in tests/test_oracle_golden.py."""
    from industrial_nnmpc_2021_b200.linearMPC import DenseQPRegulator
    g = golden_formulation
    kw = dict(A=g["unst_A"], B=g["unst_B"], Q=np.eye(2), R=np.eye(1), M=np.zeros((2, 1)), N=6,
              ulb=-np.ones((1, 1)), uub=np.ones((1, 1)))
    reg = DenseQPRegulator(**kw)
    oreg = om.DenseQPRegulatorOracle(**kw)
    assert reg.reparameterize and oreg.reparameterize
    for k in ("A", "Q", "M", "P", "tq", "G", "Krep", "Pf"):
        assert np.allclose(getattr(reg, k), getattr(oreg, k), atol=1e-11), k
    x0 = g["unst_x0"]
    assert np.allclose(reg._get_h(x0), oreg.get_h(x0), atol=1e-13)
    for scale in (0.2, 1.0, 3.0, -4.0):
        u = reg.solve(scale * x0)
        uo = oreg.solve(scale * x0)
        assert u.shape == (6, 1)
        assert np.max(np.abs(u - uo)) <= 5e-8, (scale, u.ravel(), uo.ravel())    # the oracle's IPM stops at ~1e-8


@pytest.mark.parametrize("N", [10, 30])
def test_reparameterised_regulator_matches_oracle(torch_cuda, N):
    from industrial_nnmpc_2021_b200.linearMPC import DenseQPRegulator
    A, B = unstable_plant()
    nx, nu = B.shape
    Q, R, M = np.eye(nx), 0.5 * np.eye(nu), np.zeros((nx, nu))
    kw = dict(A=A, B=B, Q=Q, R=R, M=M, N=N, ulb=-np.ones((nu, 1)), uub=np.ones((nu, 1)))
    reg = DenseQPRegulator(**kw)
    oreg = om.DenseQPRegulatorOracle(**kw)
    assert reg.reparameterize and np.allclose(reg.P, oreg.P, atol=1e-10 * np.abs(oreg.P).max())
    assert np.allclose(reg.G, oreg.G, atol=1e-11)
    rng = np.random.default_rng(N)
    X0 = rng.standard_normal((24, nx)) * np.linspace(0.05, 2.5, 24)[:, None]
    U, info = reg.solve_batch(X0)
    assert not info["maxiter_hit"]
    V = reg.to_v(U, X0)
    nact = 0
    for i in range(X0.shape[0]):
        x0 = X0[i][:, None]
        uo, oi = oreg.solve(x0, return_info=True)
        lb, ub = -np.ones(N * nu), np.ones(N * nu)
        # certificate in the space the GPU solves in, recomputed in NumPy
        assert oq.box_kkt_residual(reg._Pu, (reg._tqu @ x0)[:, 0], U[i], lb, ub) <= KKT_TOL
        nact += int(np.sum(np.abs(np.abs(U[i]) - 1.0) < 1e-12) > 0)
        assert np.max(np.abs(U[i] - uo[:, 0])) <= U0_RTOL * max(1.0, np.abs(uo).max()), (i, np.max(np.abs(U[i] - uo[:, 0])))
        # the reference's objective, evaluated at v = T^-1 (u - S x0), equals the oracle's optimal value
        cost_v = float(0.5 * V[i] @ (oreg.P @ V[i]) + (oreg.tq @ x0)[:, 0] @ V[i])
        v_or = np.linalg.solve(reg._T, uo - reg._S @ x0)[:, 0]
        cost_o = float(0.5 * v_or @ (oreg.P @ v_or) + (oreg.tq @ x0)[:, 0] @ v_or)
        assert abs(cost_v - cost_o) <= COST_RTOL * max(abs(cost_o), 1e-6)
        # the reference's own constraint G v <= h holds
        assert np.all(oreg.G @ V[i][:, None] <= oreg.get_h(x0) + 1e-9)
    assert nact >= 6, nact


def test_closed_loop_with_unstable_plant(torch_cuda):
    """simulate_offline (:827-880) on an unstable plant: target selector (I - A invertible), re-parameterised
    regulator inside the continuously batched engine, both precisions."""
    from industrial_nnmpc_2021_b200.linearMPC import OfflineSimulator
    from industrial_nnmpc_2021_b200.controller_evaluation import sample_prbs_like
    A, B = unstable_plant(nx=8, nu=4, seed=9, rho=1.03)
    nx, nu = B.shape
    rng = np.random.default_rng(1)
    C = rng.standard_normal((4, nx))
    Bd = B[:, :2].copy()
    kw = dict(A=A, B=B, C=C, H=np.zeros((0, 4)), Rs=1e-3 * np.eye(nu), Qs=np.eye(4), Bd=Bd, Cd=np.zeros((4, 2)),
              usp=np.zeros((nu, 1)), uprev=np.zeros((nu, 1)), Q=0.05 * (C.T @ C) + 1e-3 * np.eye(nx), R=np.eye(nu),
              S=0.2 * np.eye(nu), ulb=-np.ones((nu, 1)), uub=np.ones((nu, 1)), N=12)
    # (tuning chosen so that the input-space Hessian stays at cond ~5e3: an aggressive Q/R ratio on this plant gives
    #  2e5, where the first-order iteration needs > 1e4 iterations - see the conditioning table in DESIGN.md)
    T, chunks = 10, 3
    sp = sample_prbs_like(num_change=8, num_steps=T * chunks, lb=-4.0 * np.ones((4, 1)), ub=4.0 * np.ones((4, 1)),
                          mean_change=4, sigma_change=1, seed=3)
    ds = sample_prbs_like(num_change=8, num_steps=T * chunks, lb=-1.5 * np.ones((2, 1)), ub=1.5 * np.ones((2, 1)),
                          mean_change=4, sigma_change=1, seed=4)
    xprior = np.zeros((nx, 1))
    sim = OfflineSimulator(**kw, xprior=xprior, setpoints=sp, disturbances=ds, num_data_gen_task=1,
                           num_process_per_task=chunks)
    assert sim.regulator.reparameterize
    oreg = om.setup_regulator(A, B, kw["Q"], kw["R"], kw["S"], 12, kw["ulb"], kw["uub"])
    assert oreg.reparameterize
    ots = om.TargetSelectorOracle(A=A, B=B, C=C, H=kw["H"], Bd=Bd, Cd=kw["Cd"], usp=kw["usp"], Rs=kw["Rs"],
                                  Qs=kw["Qs"], ulb=kw["ulb"], uub=kw["uub"])
    for precision in ("f64", "mixed"):
        sim.engine.set_precision(precision)
        res = sim.generate_batch()
        assert float(res["kkt"].max()) <= KKT_TOL and not res["maxiter_hit"]
        assert np.sum(np.abs(res["u"]) >= 1.0 - 1e-12) >= 3, "saturated inputs wanted"
        for c in range(chunks):
            od = om.simulate_offline(x0=xprior, uprev0=kw["uprev"], A=A, B=B, Bd=Bd, regulator=oreg, ulb=kw["ulb"],
                                     uub=kw["uub"], target_selector=ots, setpoints=sp[c * T:(c + 1) * T],
                                     disturbances=ds[c * T:(c + 1) * T])
            for k in ("x", "xs", "us", "u"):
                err = np.max(np.abs(res[k][c] - od[k])) / max(1.0, np.max(np.abs(od[k])))
                assert err <= 2e-6, (precision, c, k, err)     # unstable loop: oracle IPM error (1e-8) grows along the chunk
