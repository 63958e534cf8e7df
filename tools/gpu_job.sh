#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02t: output-constrained target selector tests
set -x
mkdir -p gpurun_out
T=r02t
timeout -k 10 600 python -m pytest tests/test_gpu_target_selector_outputs.py -q > gpurun_out/${T}_pytest_ts.log 2>&1
grep -v "Warning\|^  \|^$\|warnings.html" gpurun_out/${T}_pytest_ts.log | tail -40 | cut -c1-300
