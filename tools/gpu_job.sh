#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02ag: INT8 exact GEMM epilogue with its partial-sum / scale loads issued ahead of the level fold, batched loads in the
#        KKT functor: parity, full-size, network tests; bench A/B (closed loop, network)
set -x
mkdir -p gpurun_out
T=r02ag
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cdu_fullsize.py -q -x > gpurun_out/${T}_pytest.log 2>&1
tail -2 gpurun_out/${T}_pytest.log | cut -c1-300
timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab.json 2> gpurun_out/${T}_ab.err
tail -c 300 gpurun_out/${T}_ab.err; cut -c1-1200 gpurun_out/${T}_ab.json
timeout -k 10 600 python bench.py --workload nn_10m --batch 2000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_nn.json 2> gpurun_out/${T}_nn.err
tail -c 300 gpurun_out/${T}_nn.err; cut -c1-500 gpurun_out/${T}_nn.json
