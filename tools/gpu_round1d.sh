#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/probes/gemm_tiles.py > gpurun_out/gemm_tiles.txt 2>&1; cat gpurun_out/gemm_tiles.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "1024 64" "2688 32"; do
  set -- $cfg
  timeout 900 python bench.py --steps 2 --warmup 3 --traj $1 --slab $2 --no-cpu-baseline > gpurun_out/bench_d_$1_$2.json 2> gpurun_out/bench_d_$1_$2.err
  tail -3 gpurun_out/bench_d_$1_$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_d_$1_$2.json").read().strip().splitlines()[-1])
print("traj $1 slab $2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "roof", round(d["roofline"]["achieved"],2), round(d["roofline"]["frac"],3), "share", round(d["roofline"]["share_of_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
PY
done
