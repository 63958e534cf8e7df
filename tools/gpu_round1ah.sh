#!/bin/bash
# round 1ah (2 GPUs): NCCL dataset gather test, closed-loop parity with 6-level anchors, weak-scaling bench line at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv; nproc; free -g | head -2
timeout -k 10 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity.py -q -x -k "two_gpu or closed_loop or sharding or oz_int8" 2>&1 | tail -4
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_ah_n2.json 2> gpurun_out/bench_ah_n2.err
tail -5 gpurun_out/bench_ah_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_ah_n2.json").read().strip().splitlines()[-1])
print("n2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "iters", d["iterations"], "e2e", round(d["e2e"]["value"]), "gather", d["gather"])
print("   breakdown", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["time_breakdown"].items() if k!="unit"})
PY
