#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
set -x
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_cdu_fullsize.py tests/test_gpu_reparam.py -x -q -s 2>&1 | tail -40 > gpurun_out/r02a_j3.log
tail -25 gpurun_out/r02a_j3.log
timeout -k 10 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_cdu_fullsize.py --deselect tests/test_gpu_reparam.py 2>&1 | tail -8 > gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_pytest.log
# conditioning sweep of the stand-in plant (short steps: 16384 slots x 8 sim steps, 2 timed steps)
for cfg in "0.7 0.1" "2 0.1" "5 0.1" "0.7 0.01" "5 0.01"; do
  set -- $cfg
  timeout -k 10 400 python bench.py --traj 16384 --slab 8 --steps 2 --warmup 3 --no-cpu-baseline --gain-norm $1 --r-weight $2 --max-iter 20000 \
    > gpurun_out/r02a_cond_g$1_r$2.json 2> gpurun_out/r02a_cond_g$1_r$2.err
  tail -c 600 gpurun_out/r02a_cond_g$1_r$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02a_cond_g$1_r$2.json"))
    print("COND", "$1", "$2", d["conditioning"], d["iterations"], d["value"], d["time_breakdown"])
except Exception as e:
    print("COND failed", "$1", "$2", e)
PY
done
timeout -k 10 120 python tools/probes/lp_accum_error.py 2>&1 | tail -5 > gpurun_out/r02a_lp_accum.txt; cat gpurun_out/r02a_lp_accum.txt
