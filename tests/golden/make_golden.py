"""Generate golden fixtures by importing the REFERENCE's own code (run in the build container).

    python tests/golden/make_golden.py

``/root/reference`` is read-only and does not exist on the GPU box, so the vectors are committed
as small ``.npz`` files next to this script.  ``cvxopt`` / ``h5py`` / ``matplotlib`` are not
installed; they are replaced by inert stubs in ``sys.modules`` — enough to import
``lib/linearMPC.py`` and ``lib/controller_evaluation.py`` and to run everything that does not
call ``cvx.solvers.qp`` (formulation matrices, dlqr, PRBS signals, stage cost).  ``np.int`` (removed
from NumPy, used at controller_evaluation.py:28) is shimmed to ``int``.

Inputs for the fixtures come from this repo's plant builders (the reference's CSTR builder needs
casadi/mpctools, its CDU builder needs the unshipped ``CDU_Model.mat``).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    for name in ("cvxopt", "h5py", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.backends", "matplotlib.backends.backend_pdf"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.pyplot"].rcParams = {}
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.backends.backend_pdf"].PdfPages = object
    if not hasattr(np, "int"):
        np.int = int
    sys.path.insert(0, "/root/reference/lib")
    import linearMPC as ref_lmpc
    import controller_evaluation as ref_ce
    return ref_lmpc, ref_ce


def make_ts_outputs(ref):
    """Output-constrained target selector (lib/linearMPC.py:242-248, :284-288): the reference's own fixed and changing
    matrices for the small CDU stand-in with output bounds -> target_selector_outputs.npz."""
    from industrial_nnmpc_2021_b200.plants import get_cdu_problem
    prob = get_cdu_problem(Nx=24, Nu=4, Ny=8, with_scenarios=False)
    rng = np.random.default_rng(21)
    ylb = -0.4 - 0.2 * rng.uniform(size=(prob.Ny, 1))
    yub = 0.3 + 0.2 * rng.uniform(size=(prob.Ny, 1))
    ts = ref.TargetSelector(A=prob.A, B=prob.B, C=prob.C, H=prob.H, Bd=prob.Bd, Cd=prob.Cd, usp=prob.usp, Rs=prob.Rs,
                            Qs=prob.Qs, ulb=prob.ulb, uub=prob.uub, ylb=ylb, yub=yub)
    ysp = 0.5 * rng.standard_normal((prob.Ny, 1))
    d = rng.standard_normal((prob.Nd, 1))
    q, h, b = ts._setup_changing_matrices(ysp, d)
    assert ts.h is None
    np.savez_compressed(os.path.join(HERE, "target_selector_outputs.npz"), ylb=ylb, yub=yub, ysp=ysp, d=d, P=ts.P, G=ts.G,
                        f=ts.f, e=ts.e, F=ts.F, tA=ts.tA, tb=ts.tb, q=q, h=h, b=b)


def main():
    ref, ref_ce = import_reference()
    if "ts_outputs" in sys.argv[1:]:
        make_ts_outputs(ref)
        return
    make_ts_outputs(ref)
    from industrial_nnmpc_2021_b200.plants import get_cstrs_problem, get_cdu_problem

    # ---- 1. PRBS signals with the reference seeds/constants (cstrs_parameters.py:328-337,
    #         cdu_parameters.py:135-143): store strided samples + checksums.
    out = {}
    p = get_cstrs_problem(with_scenarios=False)
    par = p.extra["parameters"]
    s1 = ref_ce.sample_prbs_like(num_change=1250, num_steps=150000, lb=par["lb"]["y"] * 1.02,
                                 ub=par["ub"]["y"] * 1.02, mean_change=120, sigma_change=2, seed=1)
    s2 = ref_ce.sample_prbs_like(num_change=2500, num_steps=150000, lb=par["lb"]["p"] * 1.02,
                                 ub=par["ub"]["p"] * 1.02, mean_change=60, sigma_change=5, seed=2)
    s3 = ref_ce.sample_prbs_like(num_change=894, num_steps=357600, lb=-1.05 * np.ones((4, 1)),
                                 ub=1.05 * np.ones((4, 1)), mean_change=400, sigma_change=1, seed=1)
    s4 = ref_ce.sample_prbs_like(num_change=1788, num_steps=357600,
                                 lb=-1.05 * np.array([[5., 20, 20, 20, 20]]).T,
                                 ub=1.05 * np.array([[5., 20, 20, 20, 20]]).T,
                                 mean_change=200, sigma_change=1, seed=2)
    for k, s in (("cstrs_sp", s1), ("cstrs_dist", s2), ("cdu_sp", s3), ("cdu_dist", s4)):
        out[k + "_every997"] = s[::997]
        out[k + "_sum"] = np.array([s.sum(), np.abs(s).sum(), (s * np.arange(1, s.shape[0] + 1)[:, None]).sum()])
        out[k + "_shape"] = np.array(s.shape)
    np.savez_compressed(os.path.join(HERE, "prbs.npz"), **out)

    # ---- 2. CSTR formulation at a short horizon (N=12 keeps the fixture small) + full-N checks
    out = {}
    for tag, prob, N in (("cstrs", get_cstrs_problem(with_scenarios=False), 12),
                         ("cdu_small", get_cdu_problem(Nx=24, Nu=4, Ny=8, with_scenarios=False), 10)):
        reg = ref.LinearMPCController.setup_regulator(A=prob.A, B=prob.B, Q=prob.Q, R=prob.R,
                                                      S=prob.S, N=N, ulb=prob.ulb, uub=prob.uub)
        ts = ref.LinearMPCController.setup_target_selector(
            A=prob.A, B=prob.B, C=prob.C, H=prob.H, Bd=prob.Bd, Cd=prob.Cd, usp=prob.usp,
            Qs=prob.Qs, Rs=prob.Rs, ulb=prob.ulb, uub=prob.uub)
        rng = np.random.default_rng(7)
        x0 = rng.standard_normal((reg.Nx, 1))
        ysp = rng.standard_normal((prob.Ny, 1))
        d = rng.standard_normal((prob.Nd, 1))
        q, h, b = ts._setup_changing_matrices(ysp, d)
        aug = ref.LinearMPCController.get_augmented_matrices_for_regulator(
            prob.A, prob.B, prob.Q, prob.R, prob.S)
        ell = ref.LinearMPCController.get_updated_average_stage_cost(
            x0[:prob.Nx], x0[prob.Nx:], 0.1 * x0[:prob.Nx], 0.2 * x0[prob.Nx:],
            0.3 * x0[prob.Nx:], aug[2], aug[3], aug[4], np.array([[0.7]]), 5)
        out.update({f"{tag}_{k}": v for k, v in dict(
            N=np.array(N), A=prob.A, B=prob.B, C=prob.C, Bd=prob.Bd, Q=prob.Q, R=prob.R, S=prob.S,
            Qs=prob.Qs, Rs=prob.Rs, ulb=prob.ulb, uub=prob.uub,
            P=reg.P, tq=reg.tq, G=reg.G, tA=reg.tA, tB=reg.tB, Pf=reg.Pf, Krep=reg.Krep,
            reparam=np.array(reg.reparameterize), h=reg._get_h(x0), x0=x0,
            ts_P=ts.P, ts_G=ts.G, ts_h=ts.h, ts_tA=ts.tA, ts_tb=ts.tb, ts_q=q, ts_b=b,
            ysp=ysp, d=d, Aaug=aug[0], Baug=aug[1], Qaug=aug[2], Raug=aug[3], Maug=aug[4],
            ell=ell).items()})
    # full-horizon CSTR: only digests of P, tq (the arrays are 2.3 MB)
    prob = get_cstrs_problem(with_scenarios=False)
    reg = ref.LinearMPCController.setup_regulator(A=prob.A, B=prob.B, Q=prob.Q, R=prob.R,
                                                  S=prob.S, N=prob.N, ulb=prob.ulb, uub=prob.uub)
    rng = np.random.default_rng(11)
    probe = rng.standard_normal((reg.P.shape[0], 3))
    out["cstrs_full_P_probe"] = reg.P @ probe
    out["cstrs_full_tq_probe"] = reg.tq.T @ probe
    out["cstrs_full_P_diag"] = np.diag(reg.P).copy()
    out["cstrs_full_probe"] = probe
    out["cstrs_full_Krep"] = reg.Krep
    out["cstrs_full_Pf"] = reg.Pf
    # unstable-A case: exercises the re-parameterised branch (linearMPC.py:366-382, :476-493)
    rng = np.random.default_rng(3)
    Au = np.array([[1.05, 0.1], [0.0, 0.9]])
    Bu = np.array([[0.0], [1.0]])
    regu = ref.DenseQPRegulator(A=Au, B=Bu, Q=np.eye(2), R=np.eye(1), M=np.zeros((2, 1)), N=5,
                                ulb=-np.ones((1, 1)), uub=np.ones((1, 1)))
    x0u = np.array([[0.5], [-0.2]])
    out.update(unst_A=Au, unst_B=Bu, unst_P=regu.P, unst_tq=regu.tq, unst_G=regu.G,
               unst_h=regu._get_h(x0u), unst_x0=x0u, unst_reparam=np.array(regu.reparameterize))
    np.savez_compressed(os.path.join(HERE, "formulation.npz"), **out)

    # ---- 3. closed-loop wiring: the reference's OWN simulate_offline / _split_scenarios /
    #         _get_data_for_training run with stand-in solvers (cvxopt is not installable): the regulator and the
    #         target selector are replaced by closed-form laws that USE the mutated bounds, so the fixture pins the
    #         order of operations (:845-866), the bound shift and deviation variables (:682-689), the data rows
    #         (state BEFORE the step, :868-872), the chunking (:786-801) and the training scaling (:254-271).
    out = {}
    rng = np.random.default_rng(21)
    nx, nu, ny, nd, N, T = 5, 2, 3, 2, 4, 9
    A = 0.6 * np.eye(nx) + 0.1 * rng.standard_normal((nx, nx))
    Bm = rng.standard_normal((nx, nu))
    Bd = rng.standard_normal((nx, nd))
    stub = dict(K=0.4 * rng.standard_normal((nu, nx + nu)), Mx=rng.standard_normal((nx, ny)),
                Nx=rng.standard_normal((nx, nd)), Mu=0.3 * rng.standard_normal((nu, ny)),
                Nu=0.3 * rng.standard_normal((nu, nd)))
    ulb, uub = -0.5 * np.ones((nu, 1)), 0.4 * np.ones((nu, 1))
    sps, dss = rng.standard_normal((T, ny)), rng.standard_normal((T, nd))
    x0, up0 = rng.standard_normal((nx, 1)), 0.1 * rng.standard_normal((nu, 1))
    captured = {}
    ref.H5pyTool.save_training_data = staticmethod(lambda dictionary, filename: captured.update(
        {**dictionary, "filename": filename}))
    ref.simulate_offline(3, 1, "golden.h5py", x0, up0, A, Bm, Bd, StubRegulator(stub["K"], N), ulb, uub,
                         StubTargetSelector(stub, ulb, uub), sps, dss)
    out.update({f"sim_{k}": np.asarray(v) for k, v in captured.items() if k not in ("filename", "data_gen_time")})
    out["sim_filename"] = np.array(captured["filename"])
    out.update(sim_A=A, sim_B=Bm, sim_Bd=Bd, sim_ulb=ulb, sim_uub=uub, sim_setpoints=sps, sim_disturbances=dss,
               sim_x0=x0, sim_uprev0=up0, sim_N=np.array(N), **{f"sim_stub_{k}": v for k, v in stub.items()})
    fake = types.SimpleNamespace(num_data_gen_task=3, num_process_per_task=2)
    big_sp, big_ds = rng.standard_normal((53, ny)), rng.standard_normal((53, nd))
    sp_split, ds_split = ref.OfflineSimulator._split_scenarios(fake, setpoints=big_sp, disturbances=big_ds)
    out.update(split_sp=big_sp, split_ds=big_ds,
               split_sp_out=np.asarray([[c for c in task] for task in sp_split]),
               split_ds_out=np.asarray([[c for c in task] for task in ds_split]))
    data = dict(x=rng.standard_normal((40, nx)), uprev=rng.standard_normal((40, nu)), xs=rng.standard_normal((40, nx)),
                us=rng.standard_normal((40, nu)), u=rng.standard_normal((40, nu)))
    scaled, xscale = ref_ce._get_data_for_training(data=data, num_samples=31)
    out.update({f"train_in_{k}": v for k, v in data.items()})
    out.update({f"train_out_{k}": v for k, v in scaled.items()})
    out["train_xscale"] = xscale
    np.savez_compressed(os.path.join(HERE, "closed_loop_glue.npz"), **out)

    # ---- 4. structured network, NumPy deployment form: the reference's own NeuralNetworkController methods
    #         (controller_evaluation.py:863-892; __init__ is bypassed - it builds cvxopt objects).  LinearMPCLayers.py
    #         itself needs tensorflow, which is not installed; the NumPy form is the same function (paper eq. 8).
    out = {}
    rng = np.random.default_rng(33)
    nx, nu = 7, 3
    for tag, with_uprev in (("with", True), ("without", False)):
        dims = [2 * nx + (2 if with_uprev else 1) * nu, 16, 12, 9, nu]
        ws = []
        for i in range(len(dims) - 1):
            ws.append(rng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i]))
            if i < len(dims) - 2:
                ws.append(0.1 * rng.standard_normal(dims[i + 1]))
        ctl = object.__new__(ref_ce.NeuralNetworkController)
        ctl.regulator_weights, ctl.nnwithuprev = ws, with_uprev
        ctl.xscale = rng.uniform(0.5, 2.0, nx)[:, np.newaxis]
        ctl.ulb, ctl.uub = -0.3 * np.ones((nu, 1)), 0.4 * np.ones((nu, 1))
        cols = []
        for _ in range(6):
            x, xs = rng.standard_normal((nx, 1)), rng.standard_normal((nx, 1))
            up, us = rng.uniform(-1, 1, (nu, 1)), rng.uniform(-0.5, 0.5, (nu, 1))
            xsc, xssc = ctl._get_scaled_x_xs(x, xs)
            u = ctl._get_control_input(xsc, up, xssc, us)
            raw = ctl._get_regulator_nn_output(xsc, up, xssc, us)
            cols.append(np.concatenate([x, up, xs, us, u, raw], axis=0)[:, 0])
        out[f"nn_{tag}_cols"] = np.asarray(cols)
        out[f"nn_{tag}_xscale"] = ctl.xscale[:, 0]
        out[f"nn_{tag}_nweights"] = np.array(len(ws))
        for i, w in enumerate(ws):
            out[f"nn_{tag}_w{i}"] = w
    out.update(nn_nx=np.array(nx), nn_nu=np.array(nu), nn_ulb=-0.3 * np.ones((nu, 1)), nn_uub=0.4 * np.ones((nu, 1)))
    np.savez_compressed(os.path.join(HERE, "structured_nn.npz"), **out)

    # ---- 5. online-loop host pieces (control_law's estimator side, linearMPC.py:87-176, :606-624): the reference's
    #         own KalmanFilter / LinearPlantSimulator / setup_filter on a small disturbance-augmented model.
    out = {}
    rng = np.random.default_rng(44)
    nx, nu, ny, nd = 4, 2, 3, 2
    A = 0.7 * np.eye(nx) + 0.1 * rng.standard_normal((nx, nx))
    Bm, Cm = rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx))
    Bd, Cd = rng.standard_normal((nx, nd)), np.vstack([np.eye(nd), np.zeros((ny - nd, nd))])
    Qwx, Qwd, Rv = 0.1 * np.eye(nx), 0.05 * np.eye(nd), 0.01 * np.eye(ny)
    xprior, dprior = 0.1 * rng.standard_normal((nx, 1)), np.zeros((nd, 1))
    kf = ref.LinearMPCController.setup_filter(A=A, B=Bm, C=Cm, Bd=Bd, Cd=Cd, Qwx=Qwx, Qwd=Qwd, Rv=Rv,
                                              xprior=xprior, dprior=dprior)
    aug = ref.LinearMPCController.get_augmented_matrices_for_filter(A, Bm, Cm, Bd, Cd, Qwx, Qwd)
    np.random.seed(5)
    plant = ref.LinearPlantSimulator(A=A, B=Bm, C=Cm, Bp=Bd, Rv=Rv, sample_time=1.0, x0=xprior)
    us = rng.uniform(-1, 1, (6, nu, 1))
    ps = rng.uniform(-1, 1, (6, nd, 1))
    ys, xhats = [plant.y[0]], []
    uprev = np.zeros((nu, 1))
    for k in range(6):
        xhats.append(kf.solve(ys[-1], uprev))
        ys.append(plant.step(us[k], ps[k]))
        uprev = us[k]
    out.update(kf_A=A, kf_B=Bm, kf_C=Cm, kf_Bd=Bd, kf_Cd=Cd, kf_Qwx=Qwx, kf_Qwd=Qwd, kf_Rv=Rv, kf_xprior=xprior,
               kf_dprior=dprior, kf_L=kf.L, kf_Aaug=aug[0], kf_Baug=aug[1], kf_Caug=aug[2], kf_Qwaug=aug[3],
               kf_us=us, kf_ps=ps, kf_ys=np.asarray(ys), kf_xhats=np.asarray(xhats), kf_plant_x=np.asarray(plant.x),
               kf_seed=np.array(5))
    np.savez_compressed(os.path.join(HERE, "online_loop.npz"), **out)
    print("wrote", os.listdir(HERE))


class StubRegulator:
    """Stand-in for DenseQPRegulator in the wiring fixture: u = clip(-K x0) repeated over the horizon, clipped with
    the bounds get_control_sequence has just written into the object (linearMPC.py:685-686)."""

    def __init__(self, K, N):
        self.K, self.N = K, N
        self.ulb = self.uub = None

    def solve(self, x0):
        return np.tile(np.clip(-self.K @ x0, self.ulb, self.uub), (self.N, 1))


class StubTargetSelector:
    """Stand-in for TargetSelector: an affine map of (ysp, d) with the input part clipped to the bounds."""

    def __init__(self, m, ulb, uub):
        self.m, self.ulb, self.uub = m, ulb, uub

    def solve(self, ysp, dhats):
        return (self.m["Mx"] @ ysp + self.m["Nx"] @ dhats, np.clip(self.m["Mu"] @ ysp + self.m["Nu"] @ dhats,
                                                                  self.ulb, self.uub))


if __name__ == "__main__":
    main()
