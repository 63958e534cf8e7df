#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
set -x
mkdir -p gpurun_out
T=r02c
timeout -k 10 1500 python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest.log 2>&1
tail -15 gpurun_out/${T}_pytest.log | cut -c1-300
timeout -k 10 120 python tools/probes/lp_accum_error.py 2>&1 | tail -5 > gpurun_out/${T}_lp_accum.txt; cat gpurun_out/${T}_lp_accum.txt
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_$name.json 2> gpurun_out/${T}_ab_$name.err
  tail -c 300 gpurun_out/${T}_ab_$name.err; cat gpurun_out/${T}_ab_$name.json
}
ab m128 NNMPC_LP_TILE=m128
ab m256 NNMPC_LP_TILE=m256
ab m256_t2off NNMPC_T2_FACTOR=0
ab m256_t2_1e4 NNMPC_T2_FACTOR=10000
ab m256_tail256 NNMPC_TAIL_ROWS=256
ab m256_tail64 NNMPC_TAIL_ROWS=64
timeout -k 10 600 python bench.py --workload nn_10m --steps 3 --warmup 3 > gpurun_out/${T}_nn_10m_tc.json 2> gpurun_out/${T}_nn_10m_tc.err
tail -c 300 gpurun_out/${T}_nn_10m_tc.err; cut -c1-1200 gpurun_out/${T}_nn_10m_tc.json
ls -la gpurun_out | tail -20
