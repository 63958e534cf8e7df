#!/bin/bash
# round 1v: failed checks re-anchor from their own gradient (start anchors kept); bench-shape sweep
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding" 2>&1 | tail -4
for cfg in "4 32768 16 16384" "4 32768 16 8192" "4 65536 8 16384" "4 16384 32 16384" "6 32768 16 16384" "4 49152 16 24576"; do
  set -- $cfg
  NNMPC_CADENCE=$1 timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj $2 --slab $3 --slots $4 --no-cpu-baseline > gpurun_out/bench_v_$1_$2_$3_$4.json 2> gpurun_out/bench_v_$1_$2_$3_$4.err
  tail -3 gpurun_out/bench_v_$1_$2_$3_$4.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_v_$1_$2_$3_$4.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]; qps=$2*$3*2
print("cad $1 traj $2 slab $3 slots $4", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"]["mean"], d["iterations"]["max"], "work", w, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "setup", round(d["setup_s"],1))
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"], "rows/launch", round(w["row_iterations"]*qps/max(r["launches"],1)))
PY
done
