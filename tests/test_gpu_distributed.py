"""Two-GPU run of the sharded offline data generator (NCCL): every rank must end up with the full
dataset, bitwise equal to the single-GPU run (sharding invariance).  Skipped on boxes
with fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from industrial_nnmpc_2021_b200.plants import get_cdu_problem
        from industrial_nnmpc_2021_b200.linearMPC import OfflineSimulator
        p = get_cdu_problem(Nx=24, Nu=4, Ny=8, N=10)
        nchunks, T = 7, 12
        sim = OfflineSimulator(**p.controller_kwargs(), xprior=p.xprior, setpoints=p.setpoints[:nchunks * T],
                               disturbances=p.disturbances[:nchunks * T], num_data_gen_task=1,
                               num_process_per_task=nchunks, device=f"cuda:{rank}")
        full = sim.generate_batch()                       # sharded over both ranks + NCCL all-gather
        single = sim.generate_batch(distributed=False)    # all chunks on this GPU
        ok = all(np.array_equal(full[k].cpu().numpy(), single[k]) for k in ("x", "uprev", "xs", "us", "u"))
        q.put((rank, ok, tuple(full["x"].shape)))
    finally:
        dist.destroy_process_group()


def test_two_gpu_generation_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and res[0][2] == (7, 12, 24)
