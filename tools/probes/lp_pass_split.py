"""Where does the time of the tcgen05 Douglas-Rachford pass go?  The production kernel (EpiDelta epilogue) against the
same TMA + tcgen05.mma main loop with an epilogue that only drains TMEM, and against the production epilogue without
TMA loads / MMAs, on synthetic state (nnmpc_lp_pass_probe).  NNMPC_LIB_PATH selects a build variant
(python -m industrial_nnmpc_2021_b200.build --out libnnmpc_x.so -DNNMPC_EPI_PHASED=0 ...)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from industrial_nnmpc_2021_b200 import _lib, build   # noqa: E402

if not os.environ.get("NNMPC_LIB_PATH"):
    build.build()
L = _lib.lib()
sizes = ((16384, 4480), (8192, 4480), (16384, 8960), (16384, 540))
if len(sys.argv) > 1:
    sizes = tuple(tuple(int(v) for v in a.split("x")) for a in sys.argv[1:])
for B, n in sizes:
    ms = (C.c_float * 8)()
    _lib.check(L.nnmpc_lp_pass_probe(B, n, 10, ms), "nnmpc_lp_pass_probe")
    fl = 2.0 * n * n * B
    print(f"B={B} n={n}: full pass {ms[0]:.3f} ms ({fl / ms[0] / 1e9:.0f} TFLOP/s algorithmic), main loop only {ms[1]:.3f} ms "
          f"({fl / ms[1] / 1e9:.0f} TFLOP/s; executed MMA {2 * fl / ms[1] / 1e9:.0f} TFLOP/s), epilogue only {ms[2]:.3f} ms "
          f"({42.0 * B * n / ms[2] / 1e9:.2f} TB/s of state) -> the epilogue exposes {100 * (ms[0] - ms[1]) / ms[0]:.0f} % of the pass; "
          f"deferred second term: one-term pass {ms[3]:.3f} ms + delivery GEMM {ms[4]:.3f} ms every 4th pass = "
          f"{ms[3] + ms[4] / 4:.3f} ms per iteration ({fl / (ms[3] + ms[4] / 4) / 1e9:.0f} TFLOP/s algorithmic)")
