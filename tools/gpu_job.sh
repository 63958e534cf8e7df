#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
# r02g: two-GPU validation of the sharded generator (NCCL gather, content check) and of the horizon-sweep workload
set -x
mkdir -p gpurun_out
T=r02g
nvidia-smi -L
timeout -k 10 600 python -m pytest tests/test_gpu_distributed.py -q -s > gpurun_out/${T}_dist_pytest.log 2>&1
tail -5 gpurun_out/${T}_dist_pytest.log | cut -c1-300
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -k "structured or regulator_model" > gpurun_out/${T}_nn_pytest.log 2>&1
tail -3 gpurun_out/${T}_nn_pytest.log | cut -c1-300
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --workload horizon_sweep --horizon 140 --samples 1000000 --steps 2 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${T}_sweep2_N140.json 2> gpurun_out/${T}_sweep2_N140.err
tail -c 500 gpurun_out/${T}_sweep2_N140.err; cut -c1-900 gpurun_out/${T}_sweep2_N140.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_sweep2_N140.json')); print('GATHER', d['gather'], 'e2e', d['e2e'], 'value', d['value'])"
timeout -k 10 600 python bench.py --workload nn_10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_nn_10m.json 2> gpurun_out/${T}_nn_10m.err
tail -c 300 gpurun_out/${T}_nn_10m.err; cut -c1-400 gpurun_out/${T}_nn_10m.json
