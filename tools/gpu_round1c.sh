#!/bin/bash
set -x
mkdir -p gpurun_out
for cfg in "1024 16" "1024 32" "2688 16" "4096 16"; do
  set -- $cfg
  timeout 900 python bench.py --steps 3 --warmup 3 --traj $1 --slab $2 --no-cpu-baseline > gpurun_out/bench_c_$1_$2.json 2> gpurun_out/bench_c_$1_$2.err
  tail -3 gpurun_out/bench_c_$1_$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c_$1_$2.json").read().strip().splitlines()[-1])
print("traj $1 slab $2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "roof", round(d["roofline"]["achieved"],2), round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 3 --slab 8 --no-cpu-baseline > gpurun_out/ncu_launches_c.log 2>&1
tail -2 gpurun_out/ncu_launches_c.log
