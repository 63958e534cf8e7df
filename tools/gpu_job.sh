#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02h: full GPU test suite, bench lines of all workloads, ncu captures of the two tensor-core kernels
set -x
mkdir -p gpurun_out
T=r02h
timeout -k 10 1500 python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log | cut -c1-300
grep -n "states/s\|solves/s\|max |out" gpurun_out/${T}_pytest.log | cut -c1-200
timeout -k 10 600 python bench.py --workload nn_10m --steps 5 --warmup 3 > gpurun_out/${T}_nn_10m.json 2> gpurun_out/${T}_nn_10m.err
tail -c 300 gpurun_out/${T}_nn_10m.err; cut -c1-600 gpurun_out/${T}_nn_10m.json
timeout -k 10 900 python bench.py --workload cstr_qp_1m --steps 2 --warmup 3 > gpurun_out/${T}_cstr_qp_1m.json 2> gpurun_out/${T}_cstr_qp_1m.err
tail -c 300 gpurun_out/${T}_cstr_qp_1m.err; cut -c1-600 gpurun_out/${T}_cstr_qp_1m.json
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
tail -c 300 gpurun_out/${T}_bench_default.err; cut -c1-600 gpurun_out/${T}_bench_default.json
# ncu: launch list of a short run (shares), then the two tensor-core kernels in full
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 3000 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --traj 16384 --slab 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launches.log 2>&1
tail -2 gpurun_out/${T}_ncu_launches.log | cut -c1-300
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_kernel -s 60 -c 1 -o gpurun_out/${T}_lp_gemm \
  python bench.py --traj 16384 --slab 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_lp.log 2>&1
tail -2 gpurun_out/${T}_ncu_lp.log | cut -c1-200
timeout -k 10 600 ncu --set full --clock-control none -k regex:oz_gemm2_kernel -s 40 -c 2 -o gpurun_out/${T}_oz_gemm \
  python bench.py --traj 16384 --slab 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_oz.log 2>&1
tail -2 gpurun_out/${T}_ncu_oz.log | cut -c1-200
ls -la gpurun_out | tail -12
