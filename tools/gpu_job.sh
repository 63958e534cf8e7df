#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02r: full GPU suite, bench lines of all single-GPU workloads, launch list and ncu captures of the two tensor-core kernels
set -x
mkdir -p gpurun_out
T=r02r
timeout -k 10 1500 python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log | cut -c1-300
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
tail -c 300 gpurun_out/${T}_bench_default.err; cut -c1-400 gpurun_out/${T}_bench_default.json
timeout -k 10 600 python bench.py --workload nn_10m --steps 5 --warmup 3 > gpurun_out/${T}_nn_10m.json 2> gpurun_out/${T}_nn_10m.err
tail -c 300 gpurun_out/${T}_nn_10m.err; cut -c1-300 gpurun_out/${T}_nn_10m.json
timeout -k 10 900 python bench.py --workload cstr_qp_1m --steps 2 --warmup 3 > gpurun_out/${T}_cstr_qp_1m.json 2> gpurun_out/${T}_cstr_qp_1m.err
tail -c 300 gpurun_out/${T}_cstr_qp_1m.err; cut -c1-300 gpurun_out/${T}_cstr_qp_1m.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --traj 16384 --slab 6 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches.csv | head -32
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_kernel -s 60 -c 1 -o gpurun_out/${T}_lp_gemm \
  python bench.py --traj 16384 --slab 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_lp.log 2>&1
tail -2 gpurun_out/${T}_ncu_lp.log | cut -c1-200
