"""Full-size runs of the two batched BASELINE.json configurations that are not the bench line, checked through
size-independent properties (the oracle cannot solve a million QPs in test time):

  configs[1]  CSTRs batched linear MPC, 1 048 576 random (x0, uprev, setpoint) QPs on one B200
  configs[3]  CDU structured network (RegulatorLayerWithUprev), 10 M states

Properties: every returned QP point is feasible and certified (KKT <= 1e-8, FP64); a row's result does not depend
on the batch it is solved in; random sub-samples agree with the CPU oracle to the north-star tolerances; the
structured network returns us exactly at steady state (paper eq. 8) for all 10 M rows."""
import numpy as np
import pytest

from oracle import linear_mpc as om, nn as onn, qp as oq

pytestmark = pytest.mark.gpu

U0_RTOL, COST_RTOL, KKT_TOL, NN_TOL = 1e-6, 1e-6, 1e-8, 1e-5


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def test_million_cstr_qps(torch_cuda, cstrs_problem):
    torch = torch_cuda
    from industrial_nnmpc_2021_b200.linearMPC import LinearMPCController
    p = cstrs_problem
    B = 1 << 20
    dev = torch.device("cuda")
    ts = LinearMPCController.setup_target_selector(p.A, p.B, p.C, p.H, p.Bd, p.Cd, p.usp, p.Qs, p.Rs, p.ulb, p.uub)
    reg = LinearMPCController.setup_regulator(p.A, p.B, p.Q, p.R, p.S, p.N, p.ulb, p.uub)
    rng = np.random.default_rng(2021)
    rows = rng.integers(0, p.setpoints.shape[0], B)
    YSP = torch.tensor(p.setpoints[rows], device=dev)
    D = torch.tensor(p.disturbances[rng.integers(0, p.disturbances.shape[0], B)], device=dev)
    XS, US = ts.solve_batch(YSP, D)
    sigma = torch.tensor(rng.choice([0.02, 0.1, 0.5], B), device=dev)[:, None]
    g = torch.Generator(device=dev).manual_seed(7)
    dx = sigma * torch.randn((B, p.Nx), dtype=torch.float64, device=dev, generator=g)
    ulb, uub = torch.tensor(p.ulb.T, device=dev), torch.tensor(p.uub.T, device=dev)
    uprev = ulb + (uub - ulb) * torch.rand((B, p.Nu), dtype=torch.float64, device=dev, generator=g)
    X0 = torch.cat([dx, uprev - US], dim=1).contiguous()
    LB, UB = (ulb - US).contiguous(), (uub - US).contiguous()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    U, info = reg.solve_batch(X0, LB, UB)
    e1.record()
    torch.cuda.synchronize()
    n = p.N * p.Nu
    assert tuple(U.shape) == (B, n) and not info["maxiter_hit"]
    assert float(info["kkt"].max()) <= KKT_TOL
    Us = U.view(B, p.N, p.Nu)
    assert bool((Us >= LB[:, None, :]).all()) and bool((Us <= UB[:, None, :]).all())          # exactly feasible
    active = ((Us == LB[:, None, :]) | (Us == UB[:, None, :])).any(dim=2).any(dim=1)
    frac_active = float(active.double().mean())
    assert frac_active > 0.0                                                                    # bounds are exercised
    print(f"\n1M CSTR QPs: {B / (e0.elapsed_time(e1) * 1e-3):.3e} solves/s, iterations mean "
          f"{float(info['iters'].double().mean()):.1f} max {int(info['iters'].max())}, active bounds in {frac_active:.1%}")
    # a row's result does not depend on the batch it is solved in
    sl = slice(300000, 300777)
    U2, info2 = reg.solve_batch(X0[sl].contiguous(), LB[sl].contiguous(), UB[sl].contiguous())
    assert float((U2 - U[sl]).abs().max()) <= 1e-7 and float(info2["kkt"].max()) <= KKT_TOL
    # random sub-sample against the exact CPU oracle
    oreg = om.setup_regulator(p.A, p.B, p.Q, p.R, p.S, p.N, p.ulb, p.uub)
    box = oq.BoxQP(oreg.P)
    pick = torch.cat([torch.as_tensor(rng.integers(0, B, 24), device=dev), torch.nonzero(active)[:24, 0]])
    X0h, LBh, UBh, Uh, costh = (t[pick].cpu().numpy() for t in (X0, LB, UB, U, info["cost"]))
    for i in range(len(pick)):
        q = oreg.tq @ X0h[i][:oreg.tq.shape[1]]
        lb, ub = np.tile(LBh[i], p.N), np.tile(UBh[i], p.N)
        ue, ei = box.solve(q, lb, ub)
        assert oq.box_kkt_residual(oreg.P, q, Uh[i], lb, ub) <= KKT_TOL
        assert np.max(np.abs(Uh[i][:p.Nu] - ue[:p.Nu])) <= U0_RTOL * max(np.max(np.abs(ue[:p.Nu])), 1e-3)
        assert abs(costh[i] - ei["cost"]) <= COST_RTOL * max(abs(ei["cost"]), 1e-6)


@pytest.mark.parametrize("nn_precision,nn_tol", [("tc", 1e-6), ("f64", 1e-9)])
def test_ten_million_state_structured_network(torch_cuda, nn_precision, nn_tol):
    torch = torch_cuda
    from industrial_nnmpc_2021_b200.LinearMPCLayers import RegulatorLayerWithUprev
    nx, nu, hidden, B = 252, 32, [832, 832, 832], 10_000_000
    dev = torch.device("cuda")
    rng = np.random.default_rng(4)
    dims = [2 * nx + 2 * nu] + hidden + [nu]
    ws = []
    for i in range(len(dims) - 1):
        lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
        ws.append(rng.uniform(-lim, lim, (dims[i], dims[i + 1])))
        if i < len(dims) - 2:
            ws.append(0.1 * rng.standard_normal(dims[i + 1]))
    layer = RegulatorLayerWithUprev(layer_dims=hidden + [nu], precision=nn_precision)
    layer.set_weights(ws)
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn((B, nx), dtype=torch.float64, device=dev, generator=g)
    xs = torch.randn((B, nx), dtype=torch.float64, device=dev, generator=g)
    up = 2.0 * torch.rand((B, nu), dtype=torch.float64, device=dev, generator=g) - 1.0
    us = 2.0 * torch.rand((B, nu), dtype=torch.float64, device=dev, generator=g) - 1.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = layer([x, up, xs, us])
    e1.record()
    torch.cuda.synchronize()
    assert tuple(out.shape) == (B, nu) and bool(torch.isfinite(out).all())
    print(f"\n10M-state structured network (568-832-832-832-32, {nn_precision}): {B / (e0.elapsed_time(e1) * 1e-3):.3e} states/s")
    # steady-state invariance for every row: x = xs, uprev = us  =>  u = us exactly, whatever the weights
    inv = layer([xs, us, xs, us])
    assert bool((inv == us).all())
    del inv
    # a row's result does not depend on the batch (the forward runs in internal chunks)
    for sl in (slice(0, 1000), slice(5_000_000, 5_000_700), slice(B - 333, B)):
        part = layer([x[sl].contiguous(), up[sl].contiguous(), xs[sl].contiguous(), us[sl].contiguous()])
        assert bool((part == out[sl]).all())
    # random sub-sample against the NumPy restatement of LinearMPCLayers.py:40-61
    pick = torch.as_tensor(rng.integers(0, B, 64), device=dev)
    ins = [t[pick].cpu().numpy() for t in (x, up, xs, us)]
    ref = onn.layer_call(ws, ins, True)
    err = float(np.max(np.abs(out[pick].cpu().numpy() - ref)))
    print(f"max |out - NumPy restatement| over 64 sampled rows: {err:.2e}")
    assert err <= nn_tol <= NN_TOL
