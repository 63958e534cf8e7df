#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02p: ncu of the one-term pass and of the epilogue alone (probe), second-term cadence 12 / 16
set -x
mkdir -p gpurun_out
T=r02p
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_kernel -s 36 -c 1 -o gpurun_out/${T}_lp_oneterm \
  python tools/probes/lp_pass_split.py 16384x4480 > gpurun_out/${T}_ncu_oneterm.log 2>&1
tail -2 gpurun_out/${T}_ncu_oneterm.log | cut -c1-200
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_kernel -s 24 -c 1 -o gpurun_out/${T}_lp_epionly \
  python tools/probes/lp_pass_split.py 16384x4480 > gpurun_out/${T}_ncu_epionly.log 2>&1
tail -2 gpurun_out/${T}_ncu_epionly.log | cut -c1-200
for cfg in "NNMPC_T2_EVERY=12" "NNMPC_T2_EVERY=16"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_${tag}.json 2> gpurun_out/${T}_ab_${tag}.err
  tail -c 300 gpurun_out/${T}_ab_${tag}.err; cut -c1-1200 gpurun_out/${T}_ab_${tag}.json
done
