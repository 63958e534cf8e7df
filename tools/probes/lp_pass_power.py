"""SM clock and board power while each phase of the tensor-core pass runs alone for ~1 s (nnmpc_lp_pass_probe with
NNMPC_PROBE_ONLY): is the overlapped pass slower than either part because the board is power capped?"""
import ctypes as C
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from industrial_nnmpc_2021_b200 import _lib, build   # noqa: E402

if not os.environ.get("NNMPC_LIB_PATH"):
    build.build()
L = _lib.lib()
B, n = 16384, 4480
samples = []
stop = False


def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits",
                              "-i", "0"], capture_output=True, text=True).stdout.strip().split(",")
        try:
            samples.append((time.time(), float(out[0]), float(out[1]), out[2].strip()))
        except (ValueError, IndexError):
            pass
        time.sleep(0.05)


th = threading.Thread(target=sampler, daemon=True)
th.start()
ms = (C.c_float * 8)()
_lib.check(L.nnmpc_lp_pass_probe(B, n, 5, ms), "warm-up")
for only, name, reps in ((0, "full pass (two terms)", 1200), (1, "main loop only (two terms)", 1800), (2, "epilogue only", 2000),
                         (3, "one-term pass", 1500), (4, "second-term delivery GEMM", 2500), (0, "full pass (two terms)", 1200)):
    os.environ["NNMPC_PROBE_ONLY"] = str(only)
    time.sleep(1.0)
    t0 = time.time()
    _lib.check(L.nnmpc_lp_pass_probe(B, n, reps, ms), "nnmpc_lp_pass_probe")
    t1 = time.time()
    s = [x for x in samples if t0 + 0.4 < x[0] < t1 - 0.1]
    clk = sorted(x[1] for x in s)
    pw = sorted(x[2] for x in s)
    cap = sum(1 for x in s if x[3].lower().startswith("active"))
    med = lambda v: v[len(v) // 2] if v else float("nan")
    print(f"{name}: {ms[only]:.3f} ms per pass over {t1 - t0:.1f} s; SM clock median {med(clk):.0f} MHz (min {clk[0] if clk else 0:.0f}), "
          f"power median {med(pw):.0f} W (max {pw[-1] if pw else 0:.0f}), sw_power_cap active in {cap}/{len(s)} samples")
stop = True
