"""Host-side multi-process logic on CPU (gloo, world_size 2): contiguous sharding of trajectory
chunks and the dataset all-gather restore the reference's (task, process) concatenation order
(/root/reference/lib/controller_evaluation.py:281-292).  The per-rank "engine" here is a cheap
deterministic stand-in; the CUDA engine's own invariance to batch composition is a GPU test."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_order():
    from industrial_nnmpc_2021_b200.distributed import shard_bounds, shard_counts
    for n in (0, 1, 7, 8, 149, 1024):
        for w in (1, 2, 3, 8):
            blocks = [shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == shard_counts(n, w)


def _fake_engine(sp, ds):
    """Stand-in closed loop: each chunk's rows depend only on that chunk's inputs."""
    x = np.cumsum(sp[..., :3] + ds[..., :1], axis=1)
    return dict(x=x, uprev=0.5 * x[..., :2], xs=2.0 * x, us=x[..., :2] - 1.0, u=x[..., :2] + ds[..., :2])


def _worker(rank, world, port, nchunks, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from industrial_nnmpc_2021_b200 import distributed as d
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        sp, ds = rng.standard_normal((nchunks, 6, 4)), rng.standard_normal((nchunks, 6, 2))
        assert d.is_distributed()
        full = d.generate_sharded(_fake_engine, sp, ds, device="cpu")
        ref = _fake_engine(sp, ds)
        ok = all(np.array_equal(full[k].numpy(), ref[k]) for k in d.DATASET_KEYS)
        lo, hi = d.shard_bounds(nchunks, world, rank)
        q.put((rank, ok, hi - lo, {k: tuple(v.shape) for k, v in full.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nchunks", [8, 7, 1])
def test_sharded_generation_matches_single_process(nchunks):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nchunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert sum(n for _, _, n, _ in res) == nchunks
    assert res[0][3]["x"] == (nchunks, 6, 3)
