"""Headline-configuration parity: the FULL-size CDU closed loop (252 states, 32 inputs, 90 outputs,
horizon N = 140 -> n = 4480 decision variables; cdu_parameters.py:94-102) through the production
engine - mixed tcgen05 tiers (fp16 increments + INT8-sliced exact applies) and the all-FP64 mode -
against the CPU oracle, sample by sample, as lib/linearMPC.py:845-866 defines the loop.

Nothing here trusts the numbers the kernels report about themselves: every optimal sequence is
captured (``capture=True``) and its KKT residual, cost and first move are recomputed in NumPy with
the host copy of the condensed operators, which tests/test_condense_fullsize.py pins against a
literal roll-out of the stage costs.  Tolerances are the north-star ones.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import linear_mpc as om
from oracle import qp as oq

U0_RTOL = 1e-6
COST_RTOL = 1e-6
KKT_TOL = 1e-8
T_STEPS = 4
# chunk starts inside the PRBS scenario: 0 = start-up from the origin against the first set-point /
# disturbance levels, the others straddle level changes of the 400 / 200-step holds
STARTS = (0, 198, 399, 597, 801, 1199)


def _rel(a, b, floor=1e-3):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), floor))


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _oracle_closed_loop(p, reg_host, starts, T):
    """The oracle's closed loop on the product's condensed (P, tq), target selector in the reference's full
    (xs, us) space with Nu = 32 (Rs = 1e-6 I, Qs = diag(1e-16 I, I): the ill-conditioned CDU tuning)."""
    oreg = om.CondensedRegulatorOracle(P=reg_host.P, tq=reg_host.tq, N=p.N, Nu=p.Nu, ulb=p.ulb, uub=p.uub)
    ots = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs,
                                  ulb=p.ulb, uub=p.uub)
    # the oracle regulator works on the augmented state [x; uprev] exactly like the product (linearMPC.py:626-644)
    datas = []
    for s in starts:
        xt, upt = p.xprior, p.uprev
        rows = dict(x=[], uprev=[], xs=[], us=[], u=[], useq=[], cost=[], nact=[])
        for t in range(T):
            ysp, d = p.setpoints[s + t][:, None], p.disturbances[s + t][:, None]
            xs, us = ots.solve(ysp, d)
            oreg.ulb, oreg.uub = p.ulb - us, p.uub - us                       # :685-686
            useq_dev, info = oreg.solve(np.vstack([xt - xs, upt - us]), return_info=True)
            assert info["kkt"] <= 1e-11, info
            useq = useq_dev + np.tile(us, (p.N, 1))                           # :689
            ut = useq[:p.Nu]
            for k, v in zip(("x", "uprev", "xs", "us", "u", "useq"), (xt, upt, xs, us, ut, useq)):
                rows[k].append(v[:, 0])
            rows["cost"].append(info["cost"])
            rows["nact"].append(info["n_active"])
            xt = p.A @ xt + p.B @ ut + p.Bd @ d
            upt = ut
        datas.append({k: np.asarray(v) for k, v in rows.items()})
    return datas


def _check_against_oracle(p, reg, res, datas, tag):
    P, tq, N, nu = reg.P, reg.tq, p.N, p.Nu
    worst = dict(u0=0.0, useq=0.0, cost=0.0, kkt=0.0, x=0.0, xs=0.0, us=0.0)
    for c, od in enumerate(datas):
        for k in ("x", "xs", "us", "uprev", "u"):
            assert _rel(res[k][c], od[k]) <= U0_RTOL, (tag, c, k, _rel(res[k][c], od[k]))
        worst["x"] = max(worst["x"], _rel(res["x"][c], od["x"]))
        worst["xs"] = max(worst["xs"], _rel(res["xs"][c], od["xs"]))
        worst["us"] = max(worst["us"], _rel(res["us"][c], od["us"]))
        for t in range(od["x"].shape[0]):
            # NumPy recomputation from the GPU's own dataset row and captured sequence
            x, xs, us, up = res["x"][c, t], res["xs"][c, t], res["us"][c, t], res["uprev"][c, t]
            x0 = np.concatenate([x - xs, up - us])
            q = tq @ x0
            lb, ub = np.tile(p.ulb[:, 0] - us, N), np.tile(p.uub[:, 0] - us, N)
            z = res["useq"][c, t] - np.tile(us, N)
            assert np.all(z >= lb - 1e-15) and np.all(z <= ub + 1e-15), (tag, c, t, "infeasible")
            kkt = oq.box_kkt_residual(P, q, z, lb, ub)
            cost = float(0.5 * z @ (P @ z) + q @ z)
            assert kkt <= KKT_TOL, (tag, c, t, kkt)
            assert abs(kkt - res["kkt"][c, t]) <= 1e-11 + 1e-3 * kkt, (tag, "reported kkt", kkt, res["kkt"][c, t])
            assert abs(cost - res["cost"][c, t]) <= 1e-9 * max(abs(cost), 1.0), (tag, "reported cost", cost, res["cost"][c, t])
            assert np.array_equal(res["useq"][c, t][:nu], res["u"][c, t])
            # against the oracle's optimum
            assert abs(cost - od["cost"][t]) <= COST_RTOL * max(abs(od["cost"][t]), 1e-6), (tag, c, t, cost, od["cost"][t])
            assert _rel(res["u"][c, t], od["u"][t]) <= U0_RTOL
            assert _rel(res["useq"][c, t], od["useq"][t]) <= 10 * U0_RTOL     # whole sequence, not only the first move
            worst["u0"] = max(worst["u0"], _rel(res["u"][c, t], od["u"][t]))
            worst["useq"] = max(worst["useq"], _rel(res["useq"][c, t], od["useq"][t]))
            worst["cost"] = max(worst["cost"], abs(cost - od["cost"][t]) / max(abs(od["cost"][t]), 1e-6))
            worst["kkt"] = max(worst["kkt"], kkt)
    return worst


@pytest.fixture(scope="module")
def cdu_full(torch_cuda):
    from industrial_nnmpc_2021_b200.plants import get_cdu_problem
    from industrial_nnmpc_2021_b200.linearMPC import OfflineSimulator
    p = get_cdu_problem(N=140, Nsim=1600)
    sp = np.vstack([p.setpoints[s:s + T_STEPS] for s in STARTS])
    ds = np.vstack([p.disturbances[s:s + T_STEPS] for s in STARTS])
    sim = OfflineSimulator(**p.controller_kwargs(), xprior=p.xprior, setpoints=sp, disturbances=ds,
                           num_data_gen_task=1, num_process_per_task=len(STARTS))
    datas = _oracle_closed_loop(p, sim.regulator, STARTS, T_STEPS)
    return p, sim, sp.reshape(len(STARTS), T_STEPS, -1), ds.reshape(len(STARTS), T_STEPS, -1), datas


@pytest.mark.parametrize("precision", ["mixed-notail", "mixed-fused-notail", "mixed-every4-notail", "mixed-tail3", "mixed", "f64"])
def test_full_cdu_closed_loop_matches_oracle(cdu_full, precision):
    """n = 4480, Nu = 32, against BoxQP / the (xs, us)-space target selector:
    "mixed-notail"  every iteration on lp_gemm_kernel<EpiDelta> (tcgen05 fp16, 35 column tiles; one operator term per
                    pass, the second delivered every 8th pass by lp_gemm_kernel<EpiAddX>), anchors and KKT
                    checks on oz_gemm2_kernel (INT8 tcgen05), to the last live row;
    "mixed-fused-notail" / "mixed-every4-notail"  the same with both terms in every pass / delivery every 4th pass;
    "mixed-tail3"   the same until three trajectories are left, then the FP64 tail (the hand-over the bench runs);
    "mixed"         default tail rule: six trajectories are below it, so the FP64 tail kernels do all iterations
                    and only the KKT checks run on the INT8 tier;
    "f64"           the all-FP64 DMMA mode."""
    p, sim, sp, ds, datas = cdu_full
    nact = [int(a) for d in datas for a in d["nact"]]
    assert sum(a > 0 for a in nact) >= len(nact) // 2 and max(nact) >= 50, f"active bounds wanted, got {nact}"
    eng = sim.engine
    eng.set_precision("f64" if precision == "f64" else "mixed")
    from industrial_nnmpc_2021_b200 import _lib
    tail = {"mixed-notail": 0, "mixed-fused-notail": 0, "mixed-every4-notail": 0, "mixed-tail3": 3}.get(precision, -1)
    _lib.check(_lib.lib().nnmpc_sim_set_tail_rows(eng._handle, tail), "nnmpc_sim_set_tail_rows")
    eng.set_second_term_cadence({"mixed-fused-notail": 0, "mixed-every4-notail": 4}.get(precision, 8))
    st0 = eng.stats()
    res = eng.run(p.xprior, p.uprev, sp, ds, capture=True)
    assert not res["maxiter_hit"]
    st1 = eng.stats()
    if precision.endswith("notail"):      # the tensor-core passes did run, in the form asked for
        d1, d2, dc = (st1[k] - st0[k] for k in ("tiles_one_term", "tiles_two_terms", "tiles_second_term_delivery"))
        if precision == "mixed-fused-notail":
            assert d2 > 0 and d1 == 0 and dc == 0, (d1, d2, dc)
        else:
            assert d1 > 0 and d2 == 0 and 0 < dc <= d1, (d1, d2, dc)
    eng.set_second_term_cadence(8)
    worst = _check_against_oracle(p, sim.regulator, res, datas, precision)
    print(f"\nfull CDU ({precision}): n_active per QP {nact}; worst rel. err u0 {worst['u0']:.2e}, useq "
          f"{worst['useq']:.2e}, cost {worst['cost']:.2e}, xs {worst['xs']:.2e}, us {worst['us']:.2e}; "
          f"recomputed KKT max {worst['kkt']:.2e}; iterations {res['iters'].tolist()}")


def test_full_cdu_target_selector_nu32(cdu_full):
    """Nu = 32 target selector alone (all 32 lanes, masked Cholesky of the cond ~1e6 reduced Hessian) on 200
    scenario rows against the oracle's (xs, us)-space KKT solve, plus rows pushed onto the bounds."""
    p, sim, *_ = cdu_full
    rng = np.random.default_rng(7)
    idx = rng.choice(1600, size=150, replace=False)
    ysp, d = p.setpoints[idx].copy(), p.disturbances[idx].copy()
    # 50 more rows with set-points far outside what the inputs can reach: many active bounds
    ysp2 = p.setpoints[idx[:50]] * rng.uniform(2.0, 6.0, size=(50, 1))
    d2 = p.disturbances[idx[:50]] * rng.uniform(1.0, 2.0, size=(50, 1))
    ysp, d = np.vstack([ysp, ysp2]), np.vstack([d, d2])
    xs, us, it = sim.target_selector.solve_batch(ysp, d, return_iters=True)
    assert np.all(it > 0)
    ots = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs,
                                  ulb=p.ulb, uub=p.uub)
    nact = 0
    for i in range(ysp.shape[0]):
        (xo, uo), info = ots.solve(ysp[i][:, None], d[i][:, None], return_info=True)
        nact += info["n_active"] > 0
        assert _rel(us[i], uo[:, 0]) <= U0_RTOL, (i, _rel(us[i], uo[:, 0]))
        assert _rel(xs[i], xo[:, 0]) <= U0_RTOL, (i, _rel(xs[i], xo[:, 0]))
        # steady state and bounds hold to rounding, independent of the oracle
        assert np.max(np.abs(xs[i] - (p.A @ xs[i] + p.B @ us[i] + p.Bd @ d[i]))) <= 1e-10 * max(1.0, np.abs(xs[i]).max())
        assert np.all(us[i] >= p.ulb[:, 0] - 1e-15) and np.all(us[i] <= p.uub[:, 0] + 1e-15)
    assert nact >= 40, nact


def test_double_horizon_cdu_matches_oracle(torch_cuda):
    """BASELINE.json configs[4] shape: horizon x2 (N = 280, n = 8960), one trajectory, mixed tiers."""
    from industrial_nnmpc_2021_b200.plants import get_cdu_problem
    from industrial_nnmpc_2021_b200.linearMPC import OfflineSimulator
    T = 3
    p = get_cdu_problem(N=280, Nsim=1600)
    sp, ds = p.setpoints[399:399 + T], p.disturbances[399:399 + T]
    sim = OfflineSimulator(**p.controller_kwargs(), xprior=p.xprior, setpoints=sp, disturbances=ds,
                           num_data_gen_task=1, num_process_per_task=1)
    datas = _oracle_closed_loop(p, sim.regulator, (399,), T)
    res = sim.engine.run(p.xprior, p.uprev, sp[None], ds[None], capture=True)
    assert not res["maxiter_hit"]
    worst = _check_against_oracle(p, sim.regulator, res, datas, "N=280")
    print(f"\nCDU N=280 (n=8960): worst rel. err u0 {worst['u0']:.2e}, cost {worst['cost']:.2e}, recomputed KKT "
          f"{worst['kkt']:.2e}, n_active {datas[0]['nact'].tolist()}")
