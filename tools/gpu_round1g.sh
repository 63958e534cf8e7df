#!/bin/bash
# round 1g: first bench of the mixed-precision engine against the FP64 engine
set -x
mkdir -p gpurun_out
for cfg in "mixed 2688 16" "f64 2688 16" "mixed 8192 8"; do
  set -- $cfg
  timeout -k 10 900 python bench.py --steps 3 --warmup 3 --traj $2 --slab $3 --precision $1 --no-cpu-baseline > gpurun_out/bench_g_$1_$2.json 2> gpurun_out/bench_g_$1_$2.err
  tail -3 gpurun_out/bench_g_$1_$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_g_$1_$2.json").read().strip().splitlines()[-1])
print("$1 traj $2 slab $3", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", d["solver_work_per_qp"], "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
