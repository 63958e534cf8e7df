#!/bin/bash
# round 1s: L2-aware tile order
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "lp_split or closed_loop or sharding" 2>&1 | tail -4
for cfg in "single 4 32768 16 16384" "pair 4 32768 16 16384"; do
  set -- $cfg
  NNMPC_LP_KERNEL=$1 NNMPC_CADENCE=$2 timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj $3 --slab $4 --slots $5 --precision mixed --no-cpu-baseline > gpurun_out/bench_s_$1_$2_$3_$4_$5.json 2> gpurun_out/bench_s_$1_$2_$3_$4_$5.err
  tail -3 gpurun_out/bench_s_$1_$2_$3_$4_$5.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s_$1_$2_$3_$4_$5.json").read().strip().splitlines()[-1])
print("$1 cad $2 traj $3 slab $4 slots $5", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", d["solver_work_per_qp"], "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
for k in single; do
NNMPC_LP_KERNEL=$k timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm -s 40 -c 1 -o gpurun_out/prof_lp5_$k -f python bench.py --steps 1 --warmup 3 --traj 8192 --slab 4 --slots 8192 --precision mixed --no-cpu-baseline > gpurun_out/ncu_s_$k.log 2>&1
done
