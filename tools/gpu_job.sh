#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02y: two-GPU validation with the final kernels: sharded generator (NCCL gather, content check), default bench at N = 2
set -x
mkdir -p gpurun_out
T=r02y
nvidia-smi -L
timeout -k 10 600 python -m pytest tests/test_gpu_distributed.py -q -s > gpurun_out/${T}_dist_pytest.log 2>&1
tail -4 gpurun_out/${T}_dist_pytest.log | cut -c1-300
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -c 400 gpurun_out/${T}_bench_2gpu.err; cut -c1-400 gpurun_out/${T}_bench_2gpu.json
