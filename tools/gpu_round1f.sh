#!/bin/bash
# round 1f: first run of the tcgen05 split-operator GEMM and the mixed-precision closed loop
set -x
mkdir -p gpurun_out
run() { timeout -k 10 "$1" "${@:2}"; echo "rc=$?"; }
run 240 python -m pytest tests/test_gpu_parity.py -x -q -k "lp_split and 128-128" 2>&1 | tail -15
run 300 python -m pytest tests/test_gpu_parity.py -q -k "lp_split" 2>&1 | tail -15
run 600 python -m pytest tests/test_gpu_parity.py -q -k "closed_loop or sharding" 2>&1 | tail -25
nvidia-smi --query-gpu=name,memory.used --format=csv
