#!/bin/bash
# round 1aj: longer chunk queues per bench step (the drain of a call amortised over more work)
set -x
mkdir -p gpurun_out
one() {
  tag=$1; shift
  timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_aj_$tag.json 2> gpurun_out/bench_aj_$tag.err
  tail -3 gpurun_out/bench_aj_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_aj_$tag.json").read().strip().splitlines()[-1])
print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", round(d["iterations"]["mean"],2), d["iterations"]["max"], "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "setup", round(d["setup_s"],1))
print("   breakdown", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["time_breakdown"].items() if k!="unit"})
PY
}
one q4x16 --traj 65536 --slab 16 --slots 16384
one q8x8 --traj 131072 --slab 8 --slots 16384
