#!/bin/bash
# round 1ag: time breakdown with the four profiling channels; anchors with 6 levels; cadence 3 / 5
set -x
mkdir -p gpurun_out
one() {
  tag=$1; shift
  timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_ag_$tag.json 2> gpurun_out/bench_ag_$tag.err
  tail -3 gpurun_out/bench_ag_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ag_$tag.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", round(d["iterations"]["mean"],2), d["iterations"]["max"], "work", {k: round(v,3) for k,v in w.items()}, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
print("   breakdown", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["time_breakdown"].items() if k!="unit"})
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
}
one base
NNMPC_LIB_PATH=$PWD/industrial_nnmpc_2021_b200/csrc/libnnmpc_a5.so one a5
NNMPC_CADENCE=3 one cad3
NNMPC_CADENCE=5 one cad5
