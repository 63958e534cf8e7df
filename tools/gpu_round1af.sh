#!/bin/bash
# round 1af: validation of HEAD - all GPU tests, smoke, default bench line (+ CPU baseline), reference arm, a larger
# slot count, steady-state launch list, full ncu captures of the tcgen05 pass and the INT8 exact GEMM
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv; nproc
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout -k 10 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout -k 10 900 python bench.py > gpurun_out/bench_af_default.json 2> gpurun_out/bench_af_default.err
tail -3 gpurun_out/bench_af_default.err; cut -c1-3500 gpurun_out/bench_af_default.json
timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj 49152 --slab 16 --slots 24576 --no-cpu-baseline > gpurun_out/bench_af_24k.json 2> gpurun_out/bench_af_24k.err
tail -3 gpurun_out/bench_af_24k.err
python - <<'PY'
import json
for f in ("default","24k"):
    d=json.loads(open(f"gpurun_out/bench_af_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "e2e", round(d["e2e"]["value"]), "breakdown", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["time_breakdown"].items() if k!="unit"})
PY
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_af_ref.json 2> gpurun_out/bench_af_ref.err
tail -3 gpurun_out/bench_af_ref.err; cut -c1-300 gpurun_out/bench_af_ref.json
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 2400 --csv --log-file gpurun_out/launches_af.csv python bench.py --steps 1 --warmup 3 --traj 16384 --slab 4 --slots 16384 --no-cpu-baseline > gpurun_out/ncu_launches_af.log 2>&1
python tools/launch_summary.py gpurun_out/launches_af.csv | head -36
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_kernel -s 40 -c 1 -o gpurun_out/prof_af_lp -f python bench.py --steps 1 --warmup 3 --traj 8192 --slab 4 --slots 8192 --no-cpu-baseline > gpurun_out/ncu_af_lp.log 2>&1
ncu -i gpurun_out/prof_af_lp.ncu-rep --page raw --csv > gpurun_out/prof_af_lp_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/prof_af_lp_raw.csv 0
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm2 -s 12 -c 2 -o gpurun_out/prof_af_oz -f python bench.py --steps 1 --warmup 3 --traj 8192 --slab 4 --slots 8192 --no-cpu-baseline > gpurun_out/ncu_af_oz.log 2>&1
ncu -i gpurun_out/prof_af_oz.ncu-rep --page raw --csv > gpurun_out/prof_af_oz_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/prof_af_oz_raw.csv 0
python tools/ncu_extract.py gpurun_out/prof_af_oz_raw.csv 1
