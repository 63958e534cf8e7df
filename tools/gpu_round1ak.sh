#!/bin/bash
# round 1ak: default bench line with 65536 chunks per step (4 rounds over 16384 slots), cycled scenario slabs
set -x
mkdir -p gpurun_out
free -g | head -2
timeout -k 10 900 python bench.py > gpurun_out/bench_ak_default.json 2> gpurun_out/bench_ak_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_ak_default.json").read().strip().splitlines()[-1])
print("default", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "e2e", d["e2e"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("   breakdown", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["time_breakdown"].items() if k!="unit"})
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
