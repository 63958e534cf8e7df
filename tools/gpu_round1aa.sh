#!/bin/bash
# round 1aa: INT8 exact GEMM throughput vs row count (three kernel variants) + one full ncu capture
set -x
mkdir -p gpurun_out
for v in 2 1 0; do
  NNMPC_OZ_VARIANT=$v timeout -k 10 300 python tools/probes/oz_rates.py 2>&1 | tail -9
done
NNMPC_OZ_VARIANT=2 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm2 -s 24 -c 2 -o gpurun_out/prof_aa_oz -f python tools/probes/oz_rates.py > gpurun_out/ncu_aa_oz.log 2>&1
tail -2 gpurun_out/ncu_aa_oz.log
ncu -i gpurun_out/prof_aa_oz.ncu-rep --page raw --csv > gpurun_out/prof_aa_oz_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/prof_aa_oz_raw.csv 0
python tools/ncu_extract.py gpurun_out/prof_aa_oz_raw.csv 1
