#!/bin/bash
# round 1e: validate HEAD on a B200 — GPU parity tests, default bench line (with cpu_baseline), reference arm
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1200 python bench.py > gpurun_out/bench_e_default.json 2> gpurun_out/bench_e_default.err
tail -3 gpurun_out/bench_e_default.err; cat gpurun_out/bench_e_default.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_e_ref.json 2> gpurun_out/bench_e_ref.err
tail -3 gpurun_out/bench_e_ref.err; cat gpurun_out/bench_e_ref.json | cut -c1-1200
