#!/bin/bash
# round 1ab: CTA-pair tcgen05 pass re-measured + profiled; larger slot counts
set -x
mkdir -p gpurun_out
run() {  # tag env... -- bench args
  tag=$1; shift
  env "$@" > /dev/null 2>&1 || true
}
for cfg in "pair4 pair 4 32768 16 16384" "s32k single 4 65536 8 32768" "s32k8 single 8 65536 8 32768" "s24k6 single 6 49152 16 24576"; do
  set -- $cfg
  NNMPC_LP_KERNEL=$2 NNMPC_CADENCE=$3 timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj $4 --slab $5 --slots $6 --no-cpu-baseline > gpurun_out/bench_ab_$1.json 2> gpurun_out/bench_ab_$1.err
  tail -3 gpurun_out/bench_ab_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_$1.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("$1", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"]["mean"], "work", w, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
NNMPC_LP_KERNEL=pair timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_pair -s 40 -c 1 -o gpurun_out/prof_ab_pair -f python bench.py --steps 1 --warmup 3 --traj 8192 --slab 4 --slots 8192 --no-cpu-baseline > gpurun_out/ncu_ab_pair.log 2>&1
tail -2 gpurun_out/ncu_ab_pair.log
ncu -i gpurun_out/prof_ab_pair.ncu-rep --page raw --csv > gpurun_out/prof_ab_pair_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/prof_ab_pair_raw.csv 0
