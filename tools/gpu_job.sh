#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02ad: final state of round 2 - full GPU suite, default bench line (CPU baseline and e2e included), launch list
set -x
mkdir -p gpurun_out
T=r02ad
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log | cut -c1-300
timeout -k 10 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
tail -c 300 gpurun_out/${T}_bench_default.err; cut -c1-300 gpurun_out/${T}_bench_default.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --traj 16384 --slab 6 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches.csv | head -14
