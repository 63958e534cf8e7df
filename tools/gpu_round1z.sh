#!/bin/bash
# round 1z: steady-state launch list of the INT8-exact mixed engine + one more bench shape
set -x
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1800 --csv --log-file gpurun_out/launches_z.csv python bench.py --steps 1 --warmup 3 --traj 16384 --slab 4 --slots 16384 --no-cpu-baseline > gpurun_out/ncu_launches_z.log 2>&1
python tools/launch_summary.py gpurun_out/launches_z.csv --seq 0 | head -40
python - <<'PY'
import csv
rows=[l for l in open("gpurun_out/launches_z.csv") if not l.startswith("==")]
seq=[(r["Kernel Name"][:60], float(r["Metric Value"].replace(",",""))/ (1e3 if r["Metric Unit"]=="ns" else 1)) for r in csv.DictReader(rows) if r.get("Metric Name")=="gpu__time_duration.sum"]
# print one full loop worth of kernels in order
start=next(i for i,(k,v) in enumerate(seq) if "k_anchor_prep" in k and i>600)
for k,v in seq[start:start+70]: print(f"{v:9.1f} us  {k}")
PY
timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj 65536 --slab 8 --slots 16384 --no-cpu-baseline > gpurun_out/bench_z_65536.json 2> gpurun_out/bench_z_65536.err
tail -3 gpurun_out/bench_z_65536.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_z_65536.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("65536x8", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", w, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), r.get("fp64_equivalent"), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
