#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02z: conditioning sweep again (rows 1, 2, 4, 5 of profiles/r02_conditioning.md) with the deferred second term
set -x
mkdir -p gpurun_out
T=r02z
for cfg in "0.7 0.1" "2.0 0.1" "5.0 0.1" "5.0 0.01"; do
  set -- $cfg
  timeout -k 10 900 python bench.py --traj 16384 --slab 8 --steps 2 --warmup 3 --gain-norm $1 --r-weight $2 --max-iter 20000 --no-cpu-baseline --no-e2e \
    > gpurun_out/${T}_cond_g$1_r$2.json 2> gpurun_out/${T}_cond_g$1_r$2.err
  tail -c 300 gpurun_out/${T}_cond_g$1_r$2.err; cut -c1-900 gpurun_out/${T}_cond_g$1_r$2.json
done
