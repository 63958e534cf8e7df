#!/bin/bash
# round 1y: INT8 exact GEMM with the warp-uniform, fully unrolled MMA issue loop (variants 1: 64 columns, 2: 128 columns)
set -x
mkdir -p gpurun_out
for v in 1 2; do
  NNMPC_OZ_VARIANT=$v timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -k "oz_int8" 2>&1 | tail -3
done
for v in 1 2; do
  NNMPC_OZ_VARIANT=$v timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y_$v.json 2> gpurun_out/bench_y_$v.err
  tail -3 gpurun_out/bench_y_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_y_$v.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("variant $v", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", w, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
NNMPC_OZ_VARIANT=2 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2400 --csv --log-file gpurun_out/launches_y.csv python bench.py --steps 1 --warmup 3 --traj 16384 --slab 4 --slots 16384 --no-cpu-baseline > gpurun_out/ncu_launches_y.log 2>&1
python tools/launch_summary.py gpurun_out/launches_y.csv | head -34
