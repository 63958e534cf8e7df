"""How accurate is the fp32 TMEM accumulation of the tcgen05 fp16 GEMM on MLP-shaped data?  Decides whether the
structured network (tolerance 1e-5, lib/LinearMPCLayers.py) can run its layers on the split-fp16 path.
C = fp16(A) (T1 + T2)' / s against FP64 of the same fp16(A): what is left is the operator split (2^-22) and the
accumulation error."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from industrial_nnmpc_2021_b200 import _lib, build

build.build()
L = _lib.lib()
rng = np.random.default_rng(0)
for n in (832, 1024, 4480):
    M = 4096
    A = np.maximum(rng.standard_normal((M, n)), 0.0) * 1.5            # ReLU-like activations
    lim = np.sqrt(6.0 / (2 * n))
    Bt = rng.uniform(-lim, lim, (n, n))                                # Glorot-uniform weights
    At, Btt = torch.tensor(A, device="cuda"), torch.tensor(Bt, device="cuda")
    C = torch.empty((M, n), dtype=torch.float64, device="cuda")
    _lib.check(L.nnmpc_lp_gemm_test(M, n, n, _lib.dptr(At), _lib.dptr(Btt), float(np.abs(Bt).max()), _lib.dptr(C), 0, None), "t")
    torch.cuda.synchronize()
    A16 = A.astype(np.float16).astype(np.float64)
    ref = A16 @ Bt.T
    err = np.abs(C.cpu().numpy() - ref)
    mag = np.abs(A16) @ np.abs(Bt).T
    print(f"n={n}: max abs err {err.max():.3e}, rms {np.sqrt((err**2).mean()):.3e}, |C| rms {np.sqrt((ref**2).mean()):.3e}, "
          f"max err / sum|a||t| {np.max(err / mag):.3e}, mean signed err {np.mean(C.cpu().numpy() - ref):.3e}")
