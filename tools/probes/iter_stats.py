#!/usr/bin/env python
"""Probe: per-QP iteration counts of the closed-loop engine on the CDU workload (saved to
gpurun_out/iters_<precision>.npy for offline analysis of the load imbalance between trajectories)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from bench import _scenarios
from industrial_nnmpc_2021_b200.linearMPC import LinearMPCController, ClosedLoopEngine

B, T, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dev = torch.device("cuda", 0)
p, sp, ds = _scenarios(B, 2 * T, seed=101)
ts = LinearMPCController.setup_target_selector(p.A, p.B, p.C, p.H, p.Bd, p.Cd, p.usp, p.Qs, p.Rs, p.ulb, p.uub, device=dev)
reg = LinearMPCController.setup_regulator(p.A, p.B, p.Q, p.R, p.S, p.N, p.ulb, p.uub, device=dev)
eng = ClosedLoopEngine(reg, ts, p.A, p.B, p.Bd, precision=prec)
f64 = dict(dtype=torch.float64, device=dev)
x, up = p.xprior, p.uprev
out = []
for i in range(2):
    r = eng.run(x, up, torch.tensor(np.ascontiguousarray(sp[:, i * T:(i + 1) * T]), **f64),
                torch.tensor(np.ascontiguousarray(ds[:, i * T:(i + 1) * T]), **f64), resume=i > 0, max_iter=3000)
    x, up = r["x_final"], r["uprev_final"]
    out.append(r["iters"].cpu().numpy())
it = np.concatenate(out, axis=1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"iters_{prec}.npy"), it)
print("iters", it.shape, "mean", it.mean(), "max", it.max(), "per-traj sum (2nd slab) mean/max", it[:, T:].sum(1).mean(), it[:, T:].sum(1).max())
print(eng.stats())
