#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02ab: target selector warm-started from the previous step's target inside a trajectory
set -x
mkdir -p gpurun_out
T=r02ab
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cdu_fullsize.py tests/test_gpu_target_selector_outputs.py tests/test_gpu_reparam.py -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab.json 2> gpurun_out/${T}_ab.err
tail -c 300 gpurun_out/${T}_ab.err; cut -c1-1200 gpurun_out/${T}_ab.json
