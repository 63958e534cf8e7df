#!/bin/bash
# round 1ai: final default bench line of the round (6-level anchors, four-channel time breakdown) + matching launch list
set -x
mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/bench_ai_default.json 2> gpurun_out/bench_ai_default.err
tail -3 gpurun_out/bench_ai_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_ai_default.json").read().strip().splitlines()[-1])
print("default", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "e2e", round(d["e2e"]["value"]), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("   breakdown", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["time_breakdown"].items() if k!="unit"})
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 2400 --csv --log-file gpurun_out/launches_ai.csv python bench.py --steps 1 --warmup 3 --traj 16384 --slab 4 --slots 16384 --no-cpu-baseline > gpurun_out/ncu_launches_ai.log 2>&1
python tools/launch_summary.py gpurun_out/launches_ai.csv | head -36
