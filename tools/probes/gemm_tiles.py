"""Probe: rank CTA tile configurations of the FP64 tensor-core GEMM (nnmpc_gemm_bench) on the
shapes of the regulator-QP iteration.  Not part of the product path."""
import ctypes as C, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from industrial_nnmpc_2021_b200 import _lib
L = _lib.lib()
L.nnmpc_gemm_bench.restype = C.c_int
L.nnmpc_gemm_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                               C.POINTER(C.c_float)]
names = {0: "128x128 w2x4 st4 bk16", 1: "128x128 w4x4 st4 bk16", 2: "128x128 w2x4 st3 bk32", 5: "128x128 w4x4 st3 bk32",
         11: "128x128 w2x4 st2 bk32", 6: "64x64 w2x2 st4", 7: "32x32 w2x2 st6", 8: "64x128 w2x4 st4", 9: "32x64 w2x2 st6",
         10: "16x32 w1x4 st8"}
n = 4480
Bt = torch.randn(n, n, device="cuda", dtype=torch.float64)
out = {}
for M, cfgs in ((1024, (0, 1, 2, 5, 11, 8)), (2688, (0, 1, 2, 5)), (4096, (0, 1, 2, 5)), (16, (7, 9, 10, 6)),
                (40, (7, 9, 10, 6)), (100, (7, 9, 6, 8)), (300, (6, 8, 9, 0)), (600, (6, 8, 0, 1))):
    A = torch.randn(M, n, device="cuda", dtype=torch.float64)
    Cm = torch.empty(M, n, device="cuda", dtype=torch.float64)
    ref = A @ Bt.T
    for c in cfgs:
        ms = C.c_float()
        rc = L.nnmpc_gemm_bench(c, M, n, n, A.data_ptr(), Bt.data_ptr(), Cm.data_ptr(), 10 if M >= 600 else 30, C.byref(ms))
        _lib.check(rc, "bench")
        err = float((Cm - ref).abs().max())
        tf = 2.0 * M * n * n / (ms.value * 1e-3) / 1e12
        print(f"M={M:5d} cfg {c:2d} {names[c]:24s} {ms.value*1e3:9.1f} us  {tf:6.2f} TF/s  err {err:.1e}", flush=True)
