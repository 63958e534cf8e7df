#!/bin/bash
# round 1ad: slimmer tcgen05 epilogue (compare/select clip, fp32 clamp, hoisted index arithmetic), 8 vs 16 epilogue warps;
# ADMM penalty / relaxation sweep
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding or lp_split" 2>&1 | tail -4
one() {
  tag=$1; shift
  timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_ad_$tag.json 2> gpurun_out/bench_ad_$tag.err
  tail -3 gpurun_out/bench_ad_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ad_$tag.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", round(d["iterations"]["mean"],2), d["iterations"]["max"], "work", {k: round(v,3) for k,v in w.items()}, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
}
one w8n
NNMPC_LIB_PATH=$PWD/industrial_nnmpc_2021_b200/csrc/libnnmpc_e16.so one w16n
one r20 --rho-scale 2.0
one r28 --rho-scale 2.8
one r20a17 --rho-scale 2.0 --alpha 1.7
one r14a17 --rho-scale 1.4 --alpha 1.7
