#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02ah: final library of round 2 - full GPU suite, smoke(), short closed-loop bench
set -x
mkdir -p gpurun_out
T=r02ah
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1
tail -2 gpurun_out/${T}_pytest.log | cut -c1-300
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
tail -1 gpurun_out/${T}_smoke.log | cut -c1-300
timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab.json 2> gpurun_out/${T}_ab.err
tail -c 300 gpurun_out/${T}_ab.err; cut -c1-900 gpurun_out/${T}_ab.json
