#!/bin/bash
# first GPU pass of the round: parity tests, bench, launch list, one full ncu capture
set -x
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; tail -3 gpurun_out/bench_1024.err; cat gpurun_out/bench_1024.json
timeout 600 python bench.py --steps 3 --warmup 3 --traj 256 --no-cpu-baseline > gpurun_out/bench_256.json 2>&1; cat gpurun_out/bench_256.json
timeout 600 python bench.py --steps 3 --warmup 3 --traj 4096 --slab 2 --no-cpu-baseline > gpurun_out/bench_4096.json 2>&1; cat gpurun_out/bench_4096.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 3 --slab 2 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 60 -c 4 -o gpurun_out/prof_r1a python bench.py --steps 1 --warmup 3 --slab 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
