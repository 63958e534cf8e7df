#!/bin/bash
# round 1an: full ncu capture of the INT8 exact GEMM at 2048 rows (both level windows) in the standalone probe
set -x
mkdir -p gpurun_out
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:oz_gemm2 -s 48 -c 2 -o gpurun_out/prof_an_oz -f python tools/probes/oz_rates.py > gpurun_out/ncu_an_oz.log 2>&1
tail -2 gpurun_out/ncu_an_oz.log
ncu -i gpurun_out/prof_an_oz.ncu-rep --page raw --csv > gpurun_out/prof_an_oz_raw.csv 2>/dev/null
python tools/ncu_extract.py gpurun_out/prof_an_oz_raw.csv 0
python tools/ncu_extract.py gpurun_out/prof_an_oz_raw.csv 1
