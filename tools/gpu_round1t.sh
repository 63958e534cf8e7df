#!/bin/bash
# round 1t: validate HEAD (mixed tcgen05 engine as default) — GPU parity tests, default bench line, best-config
# bench line, reference arm, launch list and full ncu captures of the tcgen05 pass and the FP64 anchor GEMM
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout -k 10 900 python bench.py > gpurun_out/bench_t_default.json 2> gpurun_out/bench_t_default.err
tail -3 gpurun_out/bench_t_default.err; cut -c1-2500 gpurun_out/bench_t_default.json
NNMPC_CADENCE=4 timeout -k 10 900 python bench.py --steps 2 --warmup 3 --traj 32768 --slab 16 --slots 16384 --no-cpu-baseline > gpurun_out/bench_t_big.json 2> gpurun_out/bench_t_big.err
tail -3 gpurun_out/bench_t_big.err; cut -c1-2500 gpurun_out/bench_t_big.json
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_t_ref.json 2> gpurun_out/bench_t_ref.err
tail -3 gpurun_out/bench_t_ref.err; cut -c1-1200 gpurun_out/bench_t_ref.json
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file gpurun_out/launches_t.csv python bench.py --steps 1 --warmup 3 --traj 4096 --slab 4 --slots 4096 --no-cpu-baseline > gpurun_out/ncu_launches_t.log 2>&1
tail -2 gpurun_out/ncu_launches_t.log
python tools/launch_summary.py gpurun_out/launches_t.csv | head -30
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm -s 40 -c 1 -o gpurun_out/prof_t_lp -f python bench.py --steps 1 --warmup 3 --traj 8192 --slab 4 --slots 8192 --no-cpu-baseline > gpurun_out/ncu_t_lp.log 2>&1
tail -2 gpurun_out/ncu_t_lp.log
ncu -i gpurun_out/prof_t_lp.ncu-rep --page raw --csv > gpurun_out/prof_t_lp_raw.csv 2>/dev/null
ls -la gpurun_out
