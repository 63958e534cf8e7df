#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02v: cadence of the exact phases (2 / 3 / 5 / 6 against the default 4) with the deferred second term
set -x
mkdir -p gpurun_out
T=r02v
for cfg in "NNMPC_CADENCE=2" "NNMPC_CADENCE=3" "NNMPC_CADENCE=5" "NNMPC_CADENCE=6"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_${tag}.json 2> gpurun_out/${T}_ab_${tag}.err
  tail -c 300 gpurun_out/${T}_ab_${tag}.err; cut -c1-1200 gpurun_out/${T}_ab_${tag}.json
done
