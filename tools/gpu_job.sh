#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02aa: epilogue loads two quarters ahead (three quarters of registers in flight) against one
set -x
mkdir -p gpurun_out
T=r02aa
C=$PWD/industrial_nnmpc_2021_b200/csrc
for v in "" _d2; do
  echo "== libnnmpc$v" >> gpurun_out/${T}_lp_pass_split.txt
  NNMPC_LIB_PATH=$C/libnnmpc$v.so timeout -k 10 300 python tools/probes/lp_pass_split.py 16384x4480 8192x4480 16384x540 >> gpurun_out/${T}_lp_pass_split.txt 2>&1
done
cut -c1-700 gpurun_out/${T}_lp_pass_split.txt
NNMPC_LIB_PATH=$C/libnnmpc_d2.so timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_d2.json 2> gpurun_out/${T}_ab_d2.err
tail -c 300 gpurun_out/${T}_ab_d2.err; cut -c1-1200 gpurun_out/${T}_ab_d2.json
