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


def main():
    ref, ref_ce = import_reference()
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
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
