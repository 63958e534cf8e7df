#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02i: BASELINE configs[4] on 8 GPUs - horizon x1 / x2 / x4, dataset all-gather inside the timed step
set -x
mkdir -p gpurun_out
T=r02i
nvidia-smi -L | wc -l
run() {  # horizon samples steps
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --workload horizon_sweep --horizon $1 --samples $2 --steps $3 --warmup 3 --prof-steps 1 --no-cpu-baseline \
    > gpurun_out/${T}_sweep8_N$1.json 2> gpurun_out/${T}_sweep8_N$1.err
  tail -c 300 gpurun_out/${T}_sweep8_N$1.err; cut -c1-400 gpurun_out/${T}_sweep8_N$1.json
}
run 140 10000000 5
run 280 10000000 3
run 560 2500000 2
