#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
set -x
mkdir -p gpurun_out
T=r02e
timeout -k 10 1500 python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest.log 2>&1
tail -12 gpurun_out/${T}_pytest.log | cut -c1-300
grep -n "states/s\|solves/s\|max |out" gpurun_out/${T}_pytest.log | cut -c1-200
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_$name.json 2> gpurun_out/${T}_ab_$name.err
  tail -c 300 gpurun_out/${T}_ab_$name.err; cat gpurun_out/${T}_ab_$name.json
}
ab epi128B NNMPC_NOOP=1
timeout -k 10 600 python bench.py --workload nn_10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_nn_10m.json 2> gpurun_out/${T}_nn_10m.err
tail -c 300 gpurun_out/${T}_nn_10m.err; cut -c1-1500 gpurun_out/${T}_nn_10m.json
timeout -k 10 900 python bench.py --workload cstr_qp_1m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_cstr_qp_mixed.json 2> gpurun_out/${T}_cstr_qp_mixed.err
tail -c 300 gpurun_out/${T}_cstr_qp_mixed.err; cut -c1-1500 gpurun_out/${T}_cstr_qp_mixed.json
ls -la gpurun_out | tail -8
