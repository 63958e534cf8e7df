"""Where does the time of the tcgen05 Douglas-Rachford pass go?  The production kernel (EpiDelta epilogue) against the
same TMA + tcgen05.mma main loop with an epilogue that only drains TMEM, on synthetic state (nnmpc_lp_pass_probe)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from industrial_nnmpc_2021_b200 import _lib, build   # noqa: E402

build.build()
L = _lib.lib()
for B, n in ((16384, 4480), (8192, 4480), (16384, 8960), (16384, 540)):
    ms = (C.c_float * 2)()
    _lib.check(L.nnmpc_lp_pass_probe(B, n, 10, ms), "nnmpc_lp_pass_probe")
    fl = 2.0 * n * n * B
    print(f"B={B} n={n}: full pass {ms[0]:.3f} ms ({fl / ms[0] / 1e9:.0f} TFLOP/s algorithmic), main loop only {ms[1]:.3f} ms "
          f"({fl / ms[1] / 1e9:.0f} TFLOP/s; executed MMA {2 * fl / ms[1] / 1e9:.0f} TFLOP/s) -> the epilogue exposes "
          f"{100 * (ms[0] - ms[1]) / ms[0]:.0f} % of the pass")
