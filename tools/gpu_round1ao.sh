#!/bin/bash
# round 1ao: double-buffered operator slices in the low level window of the INT8 exact GEMM - accuracy tests,
# closed-loop parity, throughput probe
set -x
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -q -x -k "oz_int8 or closed_loop_matches or sharding" 2>&1 | tail -3
timeout -k 10 100 python tools/probes/oz_rates.py 2>&1 | tail -8
