#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
set -x
mkdir -p gpurun_out
T=r02f
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log | cut -c1-300
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_$name.json 2> gpurun_out/${T}_ab_$name.err
  tail -c 300 gpurun_out/${T}_ab_$name.err; cat gpurun_out/${T}_ab_$name.json
}
ab int64fold NNMPC_NOOP=1
timeout -k 10 600 python bench.py --workload nn_10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_nn_10m.json 2> gpurun_out/${T}_nn_10m.err
tail -c 300 gpurun_out/${T}_nn_10m.err; cut -c1-1200 gpurun_out/${T}_nn_10m.json
# BASELINE configs[4] on one GPU: horizon x1 / x2 / x4 (reduced sample counts; the 10 M-sample run needs 8 GPUs)
for cfg in "140 1000000 16384" "280 250000 16384" "560 60000 8192"; do
  set -- $cfg
  timeout -k 10 900 python bench.py --workload horizon_sweep --horizon $1 --samples $2 --traj $3 --slots $3 --steps 2 --warmup 3 --no-cpu-baseline \
     > gpurun_out/${T}_sweep_N$1.json 2> gpurun_out/${T}_sweep_N$1.err
  tail -c 400 gpurun_out/${T}_sweep_N$1.err; cut -c1-700 gpurun_out/${T}_sweep_N$1.json
done
ls -la gpurun_out | tail -8
