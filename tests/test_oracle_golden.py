"""CPU: the oracle and the product's host-side setup against fixtures generated from the
reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import linear_mpc as om
from oracle import qp as oq
from industrial_nnmpc_2021_b200 import condense
from industrial_nnmpc_2021_b200.controller_evaluation import sample_prbs_like


@pytest.mark.parametrize("tag", ["cstrs", "cdu_small"])
def test_oracle_formulation_matches_reference(golden_formulation, tag):
    g = golden_formulation
    N = int(g[f"{tag}_N"])
    reg = om.setup_regulator(g[f"{tag}_A"], g[f"{tag}_B"], g[f"{tag}_Q"], g[f"{tag}_R"], g[f"{tag}_S"], N,
                             g[f"{tag}_ulb"], g[f"{tag}_uub"])
    assert bool(g[f"{tag}_reparam"]) == reg.reparameterize
    for k in ("P", "tq", "G", "tA", "tB", "Pf", "Krep"):
        ref = g[f"{tag}_{k}"]
        assert np.allclose(getattr(reg, k), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max()), k
    assert np.array_equal(reg.get_h(g[f"{tag}_x0"]), g[f"{tag}_h"])
    aug = om.augmented_matrices_for_regulator(g[f"{tag}_A"], g[f"{tag}_B"], g[f"{tag}_Q"], g[f"{tag}_R"],
                                              g[f"{tag}_S"])
    for k, a in zip(("Aaug", "Baug", "Qaug", "Raug", "Maug"), aug):
        assert np.array_equal(a, g[f"{tag}_{k}"]), k


@pytest.mark.parametrize("tag", ["cstrs", "cdu_small"])
def test_product_recursion_matches_reference(golden_formulation, tag):
    """condense.condensed_hessian (block recursion) == the reference's dense formula."""
    g = golden_formulation
    N = int(g[f"{tag}_N"])
    P, tq = condense.condensed_hessian(g[f"{tag}_Aaug"], g[f"{tag}_Baug"], g[f"{tag}_Qaug"], g[f"{tag}_Raug"],
                                       g[f"{tag}_Maug"], g[f"{tag}_Pf"], N)
    assert np.allclose(P, g[f"{tag}_P"], rtol=0, atol=1e-13 * np.abs(P).max())
    assert np.allclose(tq, g[f"{tag}_tq"], rtol=0, atol=1e-13 * np.abs(tq).max())
    K, Pf = condense.dlqr(g[f"{tag}_Aaug"], g[f"{tag}_Baug"], g[f"{tag}_Qaug"], g[f"{tag}_Raug"], g[f"{tag}_Maug"])
    assert np.allclose(K, g[f"{tag}_Krep"], rtol=1e-10, atol=1e-12)
    assert np.allclose(Pf, g[f"{tag}_Pf"], rtol=1e-10, atol=1e-10)
    tA, tB = condense.prediction_matrices(g[f"{tag}_Aaug"], g[f"{tag}_Baug"], N)
    assert np.allclose(tA, g[f"{tag}_tA"], atol=1e-13) and np.allclose(tB, g[f"{tag}_tB"], atol=1e-13)


def test_full_horizon_cstr_hessian_digest(golden_formulation, cstrs_problem):
    """N=90 CSTR Hessian from the recursion against probes of the reference's P, tq."""
    g, p = golden_formulation, cstrs_problem
    aug = om.augmented_matrices_for_regulator(p.A, p.B, p.Q, p.R, p.S)
    K, Pf = condense.dlqr(*aug)
    assert np.allclose(K, g["cstrs_full_Krep"], rtol=1e-9, atol=1e-12)
    P, tq = condense.condensed_hessian(*aug, Pf, p.N)
    probe = g["cstrs_full_probe"]
    assert np.allclose(P @ probe, g["cstrs_full_P_probe"], rtol=1e-11, atol=1e-9)
    assert np.allclose(tq.T @ probe, g["cstrs_full_tq_probe"], rtol=1e-11, atol=1e-9)
    assert np.allclose(np.diag(P), g["cstrs_full_P_diag"], rtol=1e-12)


@pytest.mark.parametrize("tag", ["cstrs", "cdu_small"])
def test_target_selector_formulation(golden_formulation, tag):
    g = golden_formulation
    nu = g[f"{tag}_B"].shape[1]
    ny = g[f"{tag}_C"].shape[0]
    nd = g[f"{tag}_Bd"].shape[1]
    ts = om.TargetSelectorOracle(A=g[f"{tag}_A"], B=g[f"{tag}_B"], C=g[f"{tag}_C"], H=np.zeros((0, ny)),
                                 Bd=g[f"{tag}_Bd"], Cd=np.zeros((ny, nd)), usp=np.zeros((nu, 1)),
                                 Rs=g[f"{tag}_Rs"], Qs=g[f"{tag}_Qs"], ulb=g[f"{tag}_ulb"], uub=g[f"{tag}_uub"])
    for k in ("P", "G", "h", "tA", "tb"):
        assert np.allclose(getattr(ts, k), g[f"{tag}_ts_{k}"], atol=1e-14), k
    q, h, b = ts.changing(g[f"{tag}_ysp"], g[f"{tag}_d"])
    assert np.allclose(q, g[f"{tag}_ts_q"], atol=1e-13) and np.allclose(b, g[f"{tag}_ts_b"], atol=1e-13)


def test_unstable_reparameterised_formulation(golden_formulation):
    g = golden_formulation
    reg = om.DenseQPRegulatorOracle(A=g["unst_A"], B=g["unst_B"], Q=np.eye(2), R=np.eye(1), M=np.zeros((2, 1)),
                                    N=5, ulb=-np.ones((1, 1)), uub=np.ones((1, 1)))
    assert reg.reparameterize and bool(g["unst_reparam"])
    for k in ("P", "tq", "G"):
        assert np.allclose(getattr(reg, k), g[f"unst_{k}"], atol=1e-12), k
    assert np.allclose(reg.get_h(g["unst_x0"]), g["unst_h"], atol=1e-13)
    u = reg.solve(g["unst_x0"])           # general-G path of the oracle runs
    assert u.shape == (5, 1) and np.all(np.abs(u) <= 1 + 1e-9)


def test_reparameterised_host_operators_and_input_space_equivalence(golden_formulation):
    """The product's host side of the unstable-A path (linearMPC.py:366-382, :476-493, :507-509): v-space
    operators equal the reference's, and the input-space box QP the GPU solves has the image u = T v + S x0 of
    the reference's minimiser as its own (checked with the oracle's two independent exact solvers)."""
    from industrial_nnmpc_2021_b200 import condense
    from oracle import qp as oq
    g = golden_formulation
    A, B, N = g["unst_A"], g["unst_B"], 5
    Q, R, M = np.eye(2), np.eye(1), np.zeros((2, 1))
    K, Pf = condense.dlqr(A, B, Q, R, M)
    A2, Q2, M2 = condense.reparameterize(A, B, Q, R, M, K)
    assert np.max(np.abs(np.linalg.eigvals(A2))) < 1.0
    Pv, tqv = condense.condensed_hessian(A2, B, Q2, R, M2, Pf, N)
    assert np.allclose(Pv, g["unst_P"], atol=1e-12) and np.allclose(tqv, g["unst_tq"], atol=1e-12)
    Pu, tqu, T, S = condense.input_space_operators(A2, B, K, N, Pv, tqv)
    E = np.vstack([np.eye(1), -np.eye(1)])
    tE = np.kron(np.eye(N), E)
    assert np.allclose(tE @ T, g["unst_G"], atol=1e-12)
    x0 = g["unst_x0"]
    te = np.tile(np.array([[1.0], [1.0]]), (N, 1))
    assert np.allclose(te - tE @ (S @ x0), g["unst_h"], atol=1e-13)
    # the input-space Hessian is the ORIGINAL problem condensed directly (same minimiser by construction)
    Pdir, tqdir = condense.condensed_hessian(A, B, Q, R, M, Pf, N)
    assert np.allclose(Pu, Pdir, rtol=1e-10, atol=1e-10) and np.allclose(tqu, tqdir, rtol=1e-10, atol=1e-10)
    rng = np.random.default_rng(0)
    oreg = om.DenseQPRegulatorOracle(A=A, B=B, Q=Q, R=R, M=M, N=N, ulb=-np.ones((1, 1)), uub=np.ones((1, 1)))
    nact = 0
    for trial in range(12):
        x0 = x0 if trial == 0 else rng.uniform(-1.5, 1.5, (2, 1)) * (1.0 + trial / 4.0)
        u_ref = oreg.solve(x0)                                   # reference route: general-G QP in v, mapped back
        u_box, info = oq.BoxQP(Pu).solve((tqu @ x0)[:, 0], -np.ones(N), np.ones(N))
        nact += info["n_active"] > 0
        assert np.allclose(u_box, u_ref[:, 0], atol=2e-8), (trial, u_box, u_ref[:, 0])
    assert nact >= 4


def test_stage_cost(golden_formulation):
    g = golden_formulation
    for tag in ("cstrs", "cdu_small"):
        x0 = g[f"{tag}_x0"]
        nx = g[f"{tag}_A"].shape[0]
        ell = om.updated_average_stage_cost(x0[:nx], x0[nx:], 0.1 * x0[:nx], 0.2 * x0[nx:], 0.3 * x0[nx:],
                                            g[f"{tag}_Qaug"], g[f"{tag}_Raug"], g[f"{tag}_Maug"],
                                            np.array([[0.7]]), 5)
        assert np.allclose(ell, g[f"{tag}_ell"], rtol=1e-13)


def test_prbs_bit_exact(golden_prbs, cstrs_problem):
    """sample_prbs_like reproduces the reference's scenario arrays bit for bit (seeds 1 and 2)."""
    g = golden_prbs
    par = cstrs_problem.extra["parameters"]
    five = np.array([[5., 20, 20, 20, 20]]).T
    cases = dict(
        cstrs_sp=dict(num_change=1250, num_steps=150000, lb=par["lb"]["y"] * 1.02, ub=par["ub"]["y"] * 1.02,
                      mean_change=120, sigma_change=2, seed=1),
        cstrs_dist=dict(num_change=2500, num_steps=150000, lb=par["lb"]["p"] * 1.02, ub=par["ub"]["p"] * 1.02,
                        mean_change=60, sigma_change=5, seed=2),
        cdu_sp=dict(num_change=894, num_steps=357600, lb=-1.05 * np.ones((4, 1)), ub=1.05 * np.ones((4, 1)),
                    mean_change=400, sigma_change=1, seed=1),
        cdu_dist=dict(num_change=1788, num_steps=357600, lb=-1.05 * five, ub=1.05 * five, mean_change=200,
                      sigma_change=1, seed=2))
    for k, kw in cases.items():
        s = sample_prbs_like(**kw)
        assert tuple(g[k + "_shape"]) == s.shape
        assert np.array_equal(s[::997], g[k + "_every997"]), k
        chk = np.array([s.sum(), np.abs(s).sum(), (s * np.arange(1, s.shape[0] + 1)[:, None]).sum()])
        assert np.array_equal(chk, g[k + "_sum"]), k


def test_cstrs_scenarios_are_the_reference_signals(golden_prbs, cstrs_problem):
    p = cstrs_problem
    assert p.setpoints.shape == (150000, 12) and p.disturbances.shape == (150000, 5)
    z = list(p.extra["z_indices"])
    assert np.array_equal(p.setpoints[::997][:, [0, 3, 7, 8, 11]], golden_prbs["cstrs_sp_every997"][:, [0, 3, 7, 8, 11]])
    assert np.all(p.setpoints[:, 4] == 0)                       # unexp_z_indices zeroed (cstrs_parameters.py:335)
    assert np.all(p.setpoints[:, [i for i in range(12) if i not in z]] == 0)
    assert np.array_equal(p.disturbances[::997], golden_prbs["cstrs_dist_every997"])


def test_cstr_model_is_box_path(cstrs_problem):
    p = cstrs_problem
    lam = np.abs(np.linalg.eigvals(p.A)).max()
    assert 0.99 < lam < 1.0       # open-loop stable => G = tE (linearMPC.py:374-382, :481)


# ----------------------------------------------------------------------------- oracle solvers
def test_box_qp_exact_and_lqr_identity(cstrs_problem):
    p = cstrs_problem
    reg = om.setup_regulator(p.A, p.B, p.Q, p.R, p.S, 30, p.ulb, p.uub)
    rng = np.random.default_rng(0)
    n = reg.P.shape[0]
    for scale in (1e-3, 0.3, 3.0):
        x0 = scale * rng.standard_normal((reg.Nx, 1))
        u, info = reg.solve(x0, return_info=True)
        assert info["kkt"] <= 1e-11
        if info["n_active"] == 0:                                  # terminal cost = DARE => LQR law
            assert np.allclose(u[:reg.Nu], reg.Krep @ x0, atol=1e-11)
    # tiny x0 is always unconstrained
    x0 = 1e-4 * rng.standard_normal((reg.Nx, 1))
    u = reg.solve(x0)
    assert np.allclose(u[:reg.Nu], reg.Krep @ x0, rtol=1e-9, atol=1e-15)


def test_ipm_matches_exact_solver():
    rng = np.random.default_rng(1)
    n = 40
    Mx = rng.standard_normal((n, n))
    P = Mx @ Mx.T + 0.1 * np.eye(n)
    q = 3 * rng.standard_normal(n)
    lb, ub = -0.5 * np.ones(n), 0.7 * np.ones(n)
    u, info = oq.BoxQP(P).solve(q, lb, ub)
    assert info["kkt"] < 1e-12 and info["n_active"] > 0
    G = np.vstack([np.eye(n), -np.eye(n)])
    h = np.concatenate([ub, -lb])
    x1, i1 = oq.ipm_qp(P, q, G, h)
    x2, i2 = oq.ipm_qp(P, q, None, h, diagonal_G=True)
    assert i1["status"] == "optimal" and i2["status"] == "optimal"
    assert np.allclose(x1[:, 0], u, atol=1e-5) and np.allclose(x2[:, 0], u, atol=1e-5)
    x3, _ = oq.solve_general_qp(P, q, G, h)
    assert np.allclose(x3[:, 0], u, atol=1e-9)


@pytest.mark.parametrize("which", ["cstrs", "cdu_small"])
def test_target_selector_oracle_full_vs_reduced(which, cstrs_problem, cdu_small_problem):
    p = cstrs_problem if which == "cstrs" else cdu_small_problem
    ts = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs, Qs=p.Qs,
                                 ulb=p.ulb, uub=p.uub)
    for t in (0, 500, 5000, 20000):
        ysp, d = p.setpoints[t][:, None], p.disturbances[t][:, None]
        (xs, us), info = ts.solve(ysp, d, return_info=True)
        assert info["eq_res"] < 1e-10
        assert np.all(us >= p.ulb - 1e-12) and np.all(us <= p.uub + 1e-12)
        q, _, b = ts.changing(ysp, d)
        nx = p.Nx
        lb = np.concatenate([np.full(nx, -np.inf), p.ulb[:, 0]])
        ub = np.concatenate([np.full(nx, np.inf), p.uub[:, 0]])
        w2, _ = oq._eq_box_reduced(ts.P, q[:, 0], ts.tA, b[:, 0], lb, ub)
        c1 = 0.5 * np.vstack([xs, us]).T @ ts.P @ np.vstack([xs, us]) + q.T @ np.vstack([xs, us])
        c2 = 0.5 * w2.T @ ts.P @ w2 + q.T @ w2
        assert abs(c1.item() - c2.item()) <= 1e-9 * max(1.0, abs(c1.item()))


def test_post_process_data_merges_in_task_process_order(tmp_path, monkeypatch):
    """controller_evaluation.py:273-295: per-(task, process) files -> one dataset, rows in (task, process) order,
    data_gen_time averaged; the training scaling of :254-271 applied to it."""
    from industrial_nnmpc_2021_b200 import controller_evaluation as ce
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(5)
    parts = {}
    for task in range(2):
        for proc in range(3):
            d = dict(x=rng.standard_normal((4, 5)), uprev=rng.standard_normal((4, 2)), xs=rng.standard_normal((4, 5)),
                     us=rng.standard_normal((4, 2)), u=rng.standard_normal((4, 2)), data_gen_time=float(task + proc))
            parts[(task, proc)] = d
            ce.H5pyTool.save_training_data(dictionary=d, filename=f"{task}-{proc}-data.h5py")
    merged = ce._post_process_data(data_filename="data.h5py", num_data_gen_task=2, num_process_per_task=3)
    order = [(t, p) for t in range(2) for p in range(3)]
    for k in ("x", "uprev", "xs", "us", "u"):
        assert np.array_equal(merged[k], np.concatenate([parts[o][k] for o in order], axis=0))
    assert merged["data_gen_time"] == np.mean([parts[o]["data_gen_time"] for o in order])
    back = ce.H5pyTool.load_training_data("data.h5py")
    assert np.array_equal(back["x"], merged["x"]) and float(back["data_gen_time"]) == merged["data_gen_time"]
    scaled, xscale = ce._get_data_for_training(data=back, num_samples=20)
    assert np.allclose(xscale, 0.5 * (back["x"][:20].max(axis=0) - back["x"][:20].min(axis=0)))
    assert np.allclose(scaled["x"] * xscale, back["x"][:20]) and np.allclose(scaled["xs"] * xscale, back["xs"][:20])


# ------------------------------------------------------------------------------------ closed-loop wiring
@pytest.fixture(scope="module")
def golden_glue():
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "closed_loop_glue.npz")) as z:
        return {k: z[k] for k in z.files}


class _StubRegulator:
    """Same stand-in as tests/golden/make_golden.py: u = clip(-K x0) with the bounds get_control_sequence wrote."""

    def __init__(self, K, N):
        self.K, self.N = K, N
        self.ulb = self.uub = None

    def solve(self, x0):
        return np.tile(np.clip(-self.K @ x0, self.ulb, self.uub), (self.N, 1))


class _StubTargetSelector:
    def __init__(self, g, ulb, uub):
        self.g, self.ulb, self.uub = g, ulb, uub

    def solve(self, ysp, dhats):
        g = self.g
        return (g["sim_stub_Mx"] @ ysp + g["sim_stub_Nx"] @ dhats,
                np.clip(g["sim_stub_Mu"] @ ysp + g["sim_stub_Nu"] @ dhats, self.ulb, self.uub))


def test_oracle_closed_loop_wiring_matches_reference_simulate_offline(golden_glue):
    """The reference's own simulate_offline (linearMPC.py:827-880), run with stand-in solvers when the fixture was
    made, against the oracle's restatement with the same stand-ins: order of operations, bound shift, deviation
    variables, rows = state BEFORE the step - bit for bit."""
    g = golden_glue
    N = int(g["sim_N"])
    od = om.simulate_offline(x0=g["sim_x0"], uprev0=g["sim_uprev0"], A=g["sim_A"], B=g["sim_B"], Bd=g["sim_Bd"],
                             regulator=_StubRegulator(g["sim_stub_K"], N), ulb=g["sim_ulb"], uub=g["sim_uub"],
                             target_selector=_StubTargetSelector(g, g["sim_ulb"], g["sim_uub"]),
                             setpoints=g["sim_setpoints"], disturbances=g["sim_disturbances"])
    for k in ("x", "uprev", "xs", "us", "u"):
        assert np.array_equal(od[k], g[f"sim_{k}"]), k
    assert str(g["sim_filename"]) == "3-1-golden.h5py"
    # the drop-in's static glue is the same function
    from industrial_nnmpc_2021_b200.linearMPC import LinearMPCController
    reg = _StubRegulator(g["sim_stub_K"], N)
    x, up = g["sim_x"][2][:, None], g["sim_uprev"][2][:, None]
    xs, us = g["sim_xs"][2][:, None], g["sim_us"][2][:, None]
    useq = LinearMPCController.get_control_sequence(reg, x, up, xs, us, g["sim_ulb"], g["sim_uub"])
    assert np.array_equal(useq[:up.shape[0], 0], g["sim_u"][2])
    assert np.array_equal(reg.ulb, g["sim_ulb"] - us) and np.array_equal(reg.uub, g["sim_uub"] - us)


def test_split_scenarios_and_training_scaling_match_reference(golden_glue):
    """OfflineSimulator._split_scenarios (linearMPC.py:786-801: equal contiguous chunks, remainder dropped,
    [task][process]) and _get_data_for_training (controller_evaluation.py:254-271) against the reference's outputs."""
    import types
    from industrial_nnmpc_2021_b200.linearMPC import OfflineSimulator
    from industrial_nnmpc_2021_b200 import controller_evaluation as ce
    g = golden_glue
    fake = types.SimpleNamespace(num_data_gen_task=3, num_process_per_task=2)
    sp, ds = OfflineSimulator._split_scenarios(fake, setpoints=g["split_sp"], disturbances=g["split_ds"])
    assert np.array_equal(np.asarray(sp), g["split_sp_out"]) and np.array_equal(np.asarray(ds), g["split_ds_out"])
    assert g["split_sp_out"].shape[:3] == (3, 2, 8)            # 53 rows over 6 processes: 8 each, 5 dropped
    osp, ods = om.split_scenarios(g["split_sp"], g["split_ds"], 6)          # the oracle's flat (task-major) form
    assert np.array_equal(np.asarray(osp).reshape(3, 2, 8, -1), g["split_sp_out"])
    assert np.array_equal(np.asarray(ods).reshape(3, 2, 8, -1), g["split_ds_out"])
    data = {k: g[f"train_in_{k}"] for k in ("x", "uprev", "xs", "us", "u")}
    scaled, xscale = ce._get_data_for_training(data=data, num_samples=31)
    assert np.array_equal(xscale, g["train_xscale"])
    for k in ("x", "uprev", "xs", "us", "u"):
        assert np.array_equal(scaled[k], g[f"train_out_{k}"]), k


# ------------------------------------------------------------------------------------ structured network
def test_nn_oracle_matches_reference_numpy_controller():
    """oracle/nn.py against outputs of the reference's own NeuralNetworkController NumPy methods
    (controller_evaluation.py:863-892: scaling, f(x,uprev,xs,us) - f(xs,us,xs,us) + us, clip), both network forms;
    the batched layer form (LinearMPCLayers.py:40-61 restated) must agree with the column form."""
    import os
    from conftest import GOLDEN
    from oracle import nn as onn
    with np.load(os.path.join(GOLDEN, "structured_nn.npz")) as z:
        g = {k: z[k] for k in z.files}
    nx, nu = int(g["nn_nx"]), int(g["nn_nu"])
    for tag, with_uprev in (("with", True), ("without", False)):
        ws = [g[f"nn_{tag}_w{i}"] for i in range(int(g[f"nn_{tag}_nweights"]))]
        xscale = g[f"nn_{tag}_xscale"][:, None]
        cols = g[f"nn_{tag}_cols"]
        clipped = 0
        for c in cols:
            x, up, xs, us, u_ref, raw_ref = np.split(c[:, None], np.cumsum([nx, nu, nx, nu, nu]))
            u = onn.control_input(ws, x, up, xs, us, with_uprev, xscale, g["nn_ulb"], g["nn_uub"])
            assert np.array_equal(u, u_ref)
            assert np.array_equal(onn.regulator_nn_output(ws, x / xscale, up, xs / xscale, us, with_uprev), raw_ref)
            clipped += int(np.any(u_ref == g["nn_uub"]) or np.any(u_ref == g["nn_ulb"]))
            # batched layer form on the scaled inputs, before the clip
            ins = [(x / xscale).T, up.T, (xs / xscale).T, us.T] if with_uprev else [(x / xscale).T, (xs / xscale).T, us.T]
            lay = onn.layer_call(ws, ins, with_uprev)
            unclipped = onn.control_input(ws, x, up, xs, us, with_uprev, xscale)
            assert np.max(np.abs(lay.T - unclipped)) <= 1e-13
        assert clipped >= 1, "fixture should exercise the output clip"


# ------------------------------------------------------------------------------------ online-loop host pieces
def test_filter_and_plant_simulator_match_reference():
    """The estimator side of control_law (linearMPC.py:87-176, :606-624): the drop-in's host mirrors of
    setup_filter / get_augmented_matrices_for_filter / KalmanFilter.solve / LinearPlantSimulator.step against a
    run of the reference's own classes (same legacy NumPy seed for the measurement noise)."""
    import os
    from conftest import GOLDEN
    from industrial_nnmpc_2021_b200.linearMPC import LinearMPCController, LinearPlantSimulator
    with np.load(os.path.join(GOLDEN, "online_loop.npz")) as z:
        g = {k: z[k] for k in z.files}
    aug = LinearMPCController.get_augmented_matrices_for_filter(g["kf_A"], g["kf_B"], g["kf_C"], g["kf_Bd"], g["kf_Cd"],
                                                                g["kf_Qwx"], g["kf_Qwd"])
    for a, k in zip(aug, ("kf_Aaug", "kf_Baug", "kf_Caug", "kf_Qwaug")):
        assert np.array_equal(a, g[k]), k
    kf = LinearMPCController.setup_filter(g["kf_A"], g["kf_B"], g["kf_C"], g["kf_Bd"], g["kf_Cd"], g["kf_Qwx"],
                                          g["kf_Qwd"], g["kf_Rv"], g["kf_xprior"], g["kf_dprior"])
    assert np.allclose(kf.L, g["kf_L"], rtol=1e-12, atol=1e-14)
    np.random.seed(int(g["kf_seed"]))
    plant = LinearPlantSimulator(A=g["kf_A"], B=g["kf_B"], C=g["kf_C"], Bp=g["kf_Bd"], Rv=g["kf_Rv"], sample_time=1.0,
                                 x0=g["kf_xprior"])
    ys, uprev = [plant.y[0]], np.zeros((g["kf_B"].shape[1], 1))
    for k in range(g["kf_us"].shape[0]):
        xhat = kf.solve(ys[-1], uprev)
        assert np.allclose(xhat, g["kf_xhats"][k], rtol=1e-11, atol=1e-13)
        ys.append(plant.step(g["kf_us"][k], g["kf_ps"][k]))
        uprev = g["kf_us"][k]
    assert np.array_equal(np.asarray(ys), g["kf_ys"]) and np.array_equal(np.asarray(plant.x), g["kf_plant_x"])
