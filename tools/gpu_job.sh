#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02u: BASELINE configs[4] on one GPU with the deferred second term: horizon x1 / x2 / x4 (reduced sample counts)
set -x
mkdir -p gpurun_out
T=r02u
for cfg in "140 1000000 16384" "280 250000 16384" "560 60000 8192"; do
  set -- $cfg
  timeout -k 10 900 python bench.py --workload horizon_sweep --horizon $1 --samples $2 --traj $3 --slots $3 --steps 2 --warmup 3 --no-cpu-baseline \
     > gpurun_out/${T}_sweep_N$1.json 2> gpurun_out/${T}_sweep_N$1.err
  tail -c 400 gpurun_out/${T}_sweep_N$1.err; cut -c1-300 gpurun_out/${T}_sweep_N$1.json
done
