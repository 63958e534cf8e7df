#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02j: training-step tests, pass-split probe, launch list of the steady-state engine
set -x
mkdir -p gpurun_out
T=r02j
timeout -k 10 600 python -m pytest tests/test_gpu_training.py -q -s > gpurun_out/${T}_train_pytest.log 2>&1
tail -15 gpurun_out/${T}_train_pytest.log | cut -c1-300
timeout -k 10 300 python tools/probes/lp_pass_split.py > gpurun_out/${T}_lp_pass_split.txt 2>&1
cat gpurun_out/${T}_lp_pass_split.txt | cut -c1-300
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4000 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --traj 16384 --slab 6 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launches.log 2>&1
tail -2 gpurun_out/${T}_ncu_launches.log | cut -c1-300
python tools/launch_summary.py gpurun_out/${T}_launches.csv | head -30
