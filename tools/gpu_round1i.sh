#!/bin/bash
# round 1i: batch-size sweep of the mixed engine + launch list
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding" 2>&1 | tail -4
for cfg in "mixed 16384 16" "mixed 32768 8"; do
  set -- $cfg
  timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj $2 --slab $3 --precision $1 --no-cpu-baseline > gpurun_out/bench_i_$1_$2_$3.json 2> gpurun_out/bench_i_$1_$2_$3.err
  tail -3 gpurun_out/bench_i_$1_$2_$3.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_i_$1_$2_$3.json").read().strip().splitlines()[-1])
print("$1 traj $2 slab $3", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", d["solver_work_per_qp"], "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 6000 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 1 --warmup 3 --traj 8192 --slab 8 --precision mixed --no-cpu-baseline > gpurun_out/ncu_i.log 2>&1
tail -2 gpurun_out/ncu_i.log | cut -c1-300
