#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02af: last sanity run of the final library: smoke() and the parity / full-size suites
set -x
mkdir -p gpurun_out
T=r02af
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
tail -1 gpurun_out/${T}_smoke.log | cut -c1-300
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cdu_fullsize.py -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -2 gpurun_out/${T}_pytest.log | cut -c1-300
