#!/bin/bash
# round 1m: ncu --set full of the tcgen05 pass (single-CTA and CTA-pair kernels), full batch
set -x
mkdir -p gpurun_out
for k in single pair; do
NNMPC_LP_KERNEL=$k timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm -s 40 -c 2 -o gpurun_out/prof_lp_$k -f python bench.py --steps 1 --warmup 3 --traj 8192 --slab 4 --slots 8192 --precision mixed --no-cpu-baseline > gpurun_out/ncu_m_$k.log 2>&1
tail -2 gpurun_out/ncu_m_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
