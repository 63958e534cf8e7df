#!/bin/bash
# The GPU job of the current development step (overwritten per step; results land in gpurun_out/, the ones worth
# keeping are copied to profiles/).
set -x
mkdir -p gpurun_out
T=r02b
timeout -k 10 900 python -m pytest tests/test_gpu_cdu_fullsize.py tests/test_gpu_reparam.py tests/test_gpu_online.py -q -s > gpurun_out/${T}_newtests.log 2>&1
tail -5 gpurun_out/${T}_newtests.log
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -k "closed_loop or sharding or chunk_queue or resume" > gpurun_out/${T}_parity.log 2>&1
tail -3 gpurun_out/${T}_parity.log
timeout -k 10 120 python tools/probes/lp_accum_error.py 2>&1 | tail -5 > gpurun_out/${T}_lp_accum.txt; cat gpurun_out/${T}_lp_accum.txt
# A/B of the tensor-core pass: one-term tiles (factor 0 = off), CTA-pair kernel, 16 epilogue warps
ab() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 10 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ab_$name.json 2> gpurun_out/${T}_ab_$name.err
  tail -c 300 gpurun_out/${T}_ab_$name.err; cat gpurun_out/${T}_ab_$name.json
}
E16=$PWD/industrial_nnmpc_2021_b200/csrc/libnnmpc_e16.so
ab t2off NNMPC_T2_FACTOR=0
ab t2_100 NNMPC_T2_FACTOR=100
ab t2_1000 NNMPC_T2_FACTOR=1000
ab pair8 NNMPC_LP_KERNEL=pair NNMPC_T2_FACTOR=100
ab pair16 NNMPC_LP_KERNEL=pair NNMPC_T2_FACTOR=100 NNMPC_LIB_PATH=$E16
ab single16 NNMPC_T2_FACTOR=100 NNMPC_LIB_PATH=$E16
# first lines of the other BASELINE workloads (FP64 DMMA paths)
timeout -k 10 600 python bench.py --workload cstr_qp_1m --steps 2 --warmup 3 > gpurun_out/${T}_cstr_qp_1m.json 2> gpurun_out/${T}_cstr_qp_1m.err
tail -c 300 gpurun_out/${T}_cstr_qp_1m.err; cut -c1-1500 gpurun_out/${T}_cstr_qp_1m.json
timeout -k 10 600 python bench.py --workload nn_10m --steps 3 --warmup 3 > gpurun_out/${T}_nn_10m.json 2> gpurun_out/${T}_nn_10m.err
tail -c 300 gpurun_out/${T}_nn_10m.err; cut -c1-1500 gpurun_out/${T}_nn_10m.json
# ncu: the tensor-core pass in full (source-level stall reasons), then the launch list of a short run
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lp_gemm_kernel -s 60 -c 2 -o gpurun_out/${T}_lp_gemm \
  python bench.py --traj 16384 --slab 4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log
ls -la gpurun_out | tail -30
