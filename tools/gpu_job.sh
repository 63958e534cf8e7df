#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02n: clocks and power of the pass phases; parity tests after the residual-mask fix
set -x
mkdir -p gpurun_out
T=r02n
timeout -k 10 300 python tools/probes/lp_pass_power.py > gpurun_out/${T}_lp_pass_power.txt 2>&1
cat gpurun_out/${T}_lp_pass_power.txt | cut -c1-300
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cdu_fullsize.py -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
