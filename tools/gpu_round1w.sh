#!/bin/bash
# round 1w: FP64-accurate anchors / KKT checks on the INT8 tcgen05 tensor cores (oz_gemm.cuh)
set -x
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -k "oz_int8" 2>&1 | tail -8
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding" 2>&1 | tail -6
for mode in int8 dmma; do
  NNMPC_EXACT_GEMM=$mode timeout -k 10 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w_$mode.json 2> gpurun_out/bench_w_$mode.err
  tail -3 gpurun_out/bench_w_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_w_$mode.json").read().strip().splitlines()[-1])
w=d["solver_work_per_qp"]
print("$mode", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "work", w, "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k in ("roofline","roofline_second_kernel"):
    r=d.get(k)
    if r: print("  ", k, r["kernel"][:40], "ach", round(r["achieved"],2), "frac", round(r["frac"],3), "share", round(r["share_of_step"],3), "avg_ms", round(r["avg_launch_ms"],3), "launches", r["launches"])
PY
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm -s 30 -c 2 -o gpurun_out/prof_w_oz -f python bench.py --steps 1 --warmup 3 --traj 16384 --slab 4 --slots 16384 --no-cpu-baseline > gpurun_out/ncu_w_oz.log 2>&1
tail -2 gpurun_out/ncu_w_oz.log
ncu -i gpurun_out/prof_w_oz.ncu-rep --page raw --csv > gpurun_out/prof_w_oz_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/prof_w_oz_raw.csv 0
python tools/ncu_extract.py gpurun_out/prof_w_oz_raw.csv 1
