#!/bin/bash
# round 1x: INT8-sliced exact GEMM kernel variants (0: 64 columns/double-buffered sets, 1: 64 columns/JIT slices,
# 2: 128 columns, two level windows); anchors with 7 levels
set -x
mkdir -p gpurun_out
for v in 1 2; do
  NNMPC_OZ_VARIANT=$v timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -k "oz_int8" 2>&1 | tail -4
done
NNMPC_OZ_VARIANT=2 timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding" 2>&1 | tail -4
for v in 0 1 2; do
  NNMPC_OZ_VARIANT=$v timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x_$v.json 2> gpurun_out/bench_x_$v.err
  tail -3 gpurun_out/bench_x_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_x_$v.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("variant $v", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", w, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
