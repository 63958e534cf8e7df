"""Probe: throughput of the INT8-sliced FP64-accurate GEMM (nnmpc_oz_gemm_bench) against the row count, at the CDU
operator size, next to cuBLAS DGEMM on the same shapes.  NNMPC_OZ_VARIANT picks the kernel.  Not on the product path."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from industrial_nnmpc_2021_b200 import _lib
L = _lib.lib()
L.nnmpc_oz_gemm_bench.restype = C.c_int
L.nnmpc_oz_gemm_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
n = 4480
Bt = torch.randn(n, n, device="cuda", dtype=torch.float64)
print("variant", os.environ.get("NNMPC_OZ_VARIANT", "2"))
for M in (128, 512, 1024, 2048, 4096, 8192, 16384):
    A = torch.randn(M, n, device="cuda", dtype=torch.float64)
    Cm = torch.empty(M, n, device="cuda", dtype=torch.float64)
    ms = (C.c_float * 2)()
    _lib.check(L.nnmpc_oz_gemm_bench(M, n, n, A.data_ptr(), Bt.data_ptr(), Cm.data_ptr(), 5, ms), "oz bench")
    ref = A @ Bt.T
    err = float((Cm - ref).abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.matmul(A, Bt.T, out=ref)
    e1.record(); torch.cuda.synchronize()
    dg = e0.elapsed_time(e1) / 5
    fl = 2.0 * M * n * n
    print(f"M={M:6d} slice {ms[0]*1e3:8.1f} us  gemm {ms[1]*1e3:9.1f} us  {fl/(ms[1]*1e-3)/1e12:7.2f} TF/s fp64-equiv "
          f"({36*fl/(ms[1]*1e-3)/1e12:7.1f} TOP/s int8)  cuBLAS dgemm {dg*1e3:9.1f} us {fl/(dg*1e-3)/1e12:6.2f} TF/s  err {err:.1e}", flush=True)
