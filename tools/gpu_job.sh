#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02q: quarter-pipelined epilogue (loads one quarter ahead, across chunks): probe, parity, bench A/B (cadence 4 / 8)
set -x
mkdir -p gpurun_out
T=r02q
timeout -k 10 300 python tools/probes/lp_pass_split.py 16384x4480 16384x540 8192x4480 > gpurun_out/${T}_lp_pass_split.txt 2>&1
cut -c1-600 gpurun_out/${T}_lp_pass_split.txt
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cdu_fullsize.py -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
for cfg in "NNMPC_T2_EVERY=8" "NNMPC_T2_EVERY=8 NNMPC_CADENCE=8"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_${tag}.json 2> gpurun_out/${T}_ab_${tag}.err
  tail -c 300 gpurun_out/${T}_ab_${tag}.err; cut -c1-1200 gpurun_out/${T}_ab_${tag}.json
done
