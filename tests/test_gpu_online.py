"""Online loop on the GPU solvers: LinearMPCController.control_law / online_simulation
(lib/linearMPC.py:646-669, :703-718) against the oracle's restatement of the same loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import linear_mpc as om


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _controller_kwargs(p, N):
    nd = p.Nd
    return dict(A=p.A, B=p.B, C=p.C, H=p.H, Qwx=1e-6 * np.eye(p.Nx), Qwd=1e-2 * np.eye(nd), Rv=1e-4 * np.eye(p.Ny),
                xprior=np.zeros((p.Nx, 1)), dprior=np.zeros((nd, 1)), Rs=p.Rs, Qs=p.Qs, Bd=p.Bd, Cd=p.Cd, usp=p.usp,
                uprev=np.zeros((p.Nu, 1)), Q=p.Q, R=p.R, S=p.S, ulb=p.ulb, uub=p.uub, N=N)


def test_control_law_and_online_simulation_match_oracle(torch_cuda, cstrs_problem, tmp_path):
    from industrial_nnmpc_2021_b200.linearMPC import LinearMPCController, LinearPlantSimulator, online_simulation
    p = cstrs_problem
    N, Nsim, start = 30, 14, 29990          # the scenario window straddles a set-point / disturbance change
    kw = _controller_kwargs(p, N)
    ctrl = LinearMPCController(**kw)
    octrl = om.OnlineControllerOracle(**kw)
    sp, ds = p.setpoints[start:start + Nsim], p.disturbances[start:start + Nsim]
    np.random.seed(7)
    plant = LinearPlantSimulator(A=p.A, B=p.B, C=p.C, Bp=p.Bd, Rv=kw["Rv"], sample_time=10.0, x0=np.zeros((p.Nx, 1)))
    y0 = plant.y[0]
    online_simulation(plant, ctrl, setpoints=sp, disturbances=ds, Nsim=Nsim, stdout_filename=str(tmp_path / "log.txt"))
    assert len(plant.u) == Nsim and len(ctrl.computation_times) == Nsim and len(ctrl.average_stage_costs) == Nsim + 1
    assert (tmp_path / "log.txt").read_text().count("Simulation Step:") == Nsim
    # the oracle loop on a plant replaying the same measurement noise
    np.random.seed(7)
    plant2 = LinearPlantSimulator(A=p.A, B=p.B, C=p.C, Bp=p.Bd, Rv=kw["Rv"], sample_time=10.0, x0=np.zeros((p.Nx, 1)))
    assert np.array_equal(plant2.y[0], y0)
    uo = om.online_simulation(plant2.step, plant2.y[0], octrl, sp, ds, Nsim)
    ug = np.asarray(plant.u)[:, :, 0]
    assert np.max(np.abs(ug - uo)) <= 1e-6 * max(1.0, np.abs(uo).max())
    assert np.max(np.abs(np.asarray(plant.x) - np.asarray(plant2.x))) <= 1e-6
    ell_g = np.asarray(ctrl.average_stage_costs).ravel()
    ell_o = np.asarray(octrl.average_stage_costs).ravel()
    assert np.allclose(ell_g, ell_o, rtol=1e-6, atol=1e-9)
    assert np.sum(np.abs(np.abs(ug) - 1.0) < 1e-9) > 0, "the window should saturate an input"
    # control_law called directly: signature (ysp (Ny,1), y (Ny,1)) -> (Nu,1), state carried in the object
    u = ctrl.control_law(sp[-1][:, None], plant.y[-1])
    assert u.shape == (p.Nu, 1) and ctrl.uprev is u
    assert np.max(np.abs(u - octrl.control_law(sp[-1][:, None], plant2.y[-1]))) <= 1e-6


@pytest.mark.parametrize("kind", ["mpc", "nn", "satdlqr"])
def test_batched_online_loop_matches_oracle(torch_cuda, cstrs_problem, kind):
    """S scenarios in lock step through nnmpc_online_run (filter GEMM, target-selector kernel, batched regulator QP /
    structured network / saturated LQR, stage-cost and plant kernels) against the one-scenario-at-a-time oracle loops,
    with the reference's noise convention (np.random.seed(seed) before every scenario)."""
    from industrial_nnmpc_2021_b200.controller_evaluation import BatchedOnlineSimulation
    from industrial_nnmpc_2021_b200.linearMPC import LinearPlantSimulator
    p = cstrs_problem
    N, T, starts = 30, 10, (29990, 0, 59990, 119990, 89995)
    kw = _controller_kwargs(p, N)
    sp = np.stack([p.setpoints[s:s + T] for s in starts])
    ds = np.stack([p.disturbances[s:s + T] for s in starts])
    extra, okw = {}, {}
    if kind == "nn":
        rng = np.random.default_rng(5)
        dims = [2 * p.Nx + 2 * p.Nu, 48, 40, p.Nu]
        ws = []
        for i in range(3):
            lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
            ws.append(rng.uniform(-lim, lim, (dims[i], dims[i + 1])))
            if i < 2:
                ws.append(0.1 * rng.standard_normal(dims[i + 1]))
        xscale = rng.uniform(0.5, 2.0, p.Nx)
        extra = dict(regulator_weights=ws, xscale=xscale, nnwithuprev=True, precision="f64")
        okw = dict(regulator_weights=ws, xscale=xscale, nnwithuprev=True)
    sim = BatchedOnlineSimulation(kind=kind, **kw, **extra)
    res = sim.run(sp, ds, seed=3)
    assert res["u"].shape == (len(starts), T, p.Nu) and res["y"].shape == (len(starts), T + 1, p.Ny)
    Oracle = {"mpc": om.OnlineControllerOracle, "nn": om.OnlineNNControllerOracle,
              "satdlqr": om.OnlineSatDlqrControllerOracle}[kind]
    sat = 0
    for c in range(len(starts)):
        octrl = Oracle(**kw, **okw)
        np.random.seed(3)
        plant = LinearPlantSimulator(A=p.A, B=p.B, C=p.C, Bp=p.Bd, Rv=kw["Rv"], sample_time=10.0, x0=np.zeros((p.Nx, 1)))
        uo = om.online_simulation(plant.step, plant.y[0], octrl, sp[c], ds[c], T)
        scale = max(1.0, np.abs(uo).max())
        assert np.max(np.abs(res["u"][c] - uo)) <= 1e-6 * scale, (kind, c, np.max(np.abs(res["u"][c] - uo)))
        assert np.max(np.abs(res["y"][c] - np.asarray(plant.y)[:, :, 0])) <= 1e-6 * max(1.0, np.abs(np.asarray(plant.y)).max())
        assert np.max(np.abs(res["x"][c] - np.asarray(plant.x)[:, :, 0])) <= 1e-6 * max(1.0, np.abs(np.asarray(plant.x)).max())
        assert np.max(np.abs(res["xhat"][c] - np.asarray(octrl.xhats)[:, :, 0])) <= 1e-6
        ell = np.asarray(octrl.average_stage_costs).ravel()[1:]
        assert np.allclose(res["average_stage_costs"][c], ell, rtol=1e-6, atol=1e-9), (kind, c)
        sat += int(np.sum(np.abs(np.abs(uo) - 1.0) < 1e-9))
    if kind != "nn":
        assert sat > 0, "some scenario should saturate an input"
    if kind == "mpc":
        assert float(res["kkt"].max()) <= 1e-8 and not res["maxiter_hit"]
    # the validation metric of the reference (controller_evaluation.py:396-397) is defined on these results
    loss = BatchedOnlineSimulation.performance_loss(res, res)
    assert loss.shape == (len(starts),) and np.all(loss == 0.0)
