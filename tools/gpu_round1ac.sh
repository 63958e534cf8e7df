#!/bin/bash
# round 1ac: k_select with 4 rows per thread; 16 epilogue warps in the tcgen05 pass; solver parameter sweep
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding" 2>&1 | tail -4
one() {
  tag=$1; shift
  timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_ac_$tag.json 2> gpurun_out/bench_ac_$tag.err
  tail -3 gpurun_out/bench_ac_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ac_$tag.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", round(d["iterations"]["mean"],2), d["iterations"]["max"], "work", {k: round(v,3) for k,v in w.items()}, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
}
one w8
NNMPC_LIB_PATH=$PWD/industrial_nnmpc_2021_b200/csrc/libnnmpc_e16.so one w16
one a17 --alpha 1.7
one a19 --alpha 1.9
one r07 --rho-scale 0.7
one r14 --rho-scale 1.4
