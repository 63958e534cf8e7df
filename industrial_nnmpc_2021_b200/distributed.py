"""Multi-GPU plumbing of the offline data generator: shard trajectory chunks over ranks, gather
the generated dataset.

The reference parallelises data generation over independent OS processes / cluster tasks, each
writing its own file, and "gathers" by concatenating the files in ``(task, process)`` order
(/root/reference/lib/linearMPC.py:786-825, lib/controller_evaluation.py:273-295).  Here the
trajectory chunks are sharded in contiguous blocks over the ranks of a ``torch.distributed`` group
(one process per GPU); there is no collective on the solve path, and one all-gather per dataset
array (NCCL over NVLink on GPUs, gloo in the CPU tests) restores the reference's order.
"""
from __future__ import annotations

import numpy as np

DATASET_KEYS = ("x", "uprev", "xs", "us", "u")


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed():
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_bounds(num_items, world_size, rank):
    """Contiguous block [lo, hi) of ``num_items`` owned by ``rank``; block sizes differ by at
    most one and concatenating the blocks in rank order restores the original order."""
    base, rem = divmod(int(num_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(num_items, world_size):
    return [shard_bounds(num_items, world_size, r)[1] - shard_bounds(num_items, world_size, r)[0]
            for r in range(world_size)]


def gather_chunks(local, num_items, group=None, keys=None):
    """All-gather per-rank dataset blocks into the full arrays, on every rank.

    ``local[k]`` is this rank's block ``(n_local, ...)`` (torch tensor on the group's device type:
    CUDA for NCCL, CPU for gloo); returns ``{k: (num_items, ...)}`` in global chunk order.  Blocks
    are padded to the largest one because ``all_gather_into_tensor`` wants equal sizes.
    """
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    counts = shard_counts(num_items, world)
    nmax = max(counts)
    out = {}
    for k in (keys or list(local)):
        t = local[k]
        if t.shape[0] != counts[dist.get_rank(group)]:
            raise ValueError(f"{k}: local block has {t.shape[0]} chunks, expected {counts[dist.get_rank(group)]}")
        if t.shape[0] < nmax:
            pad = torch.zeros((nmax - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], dim=0)
        t = t.contiguous()
        full = torch.empty((world * nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t, group=group)
        if all(c == nmax for c in counts):
            out[k] = full
        else:
            out[k] = torch.cat([full[r * nmax:r * nmax + c] for r, c in enumerate(counts)], dim=0)
    return out


def generate_sharded(run_local, setpoints, disturbances, group=None, device=None, keys=DATASET_KEYS):
    """Shard ``(num_chunks, L, .)`` scenario arrays over the ranks, run ``run_local(sp, ds)`` on
    this rank's block (it returns a dict of ``(n_local, L, .)`` tensors/arrays) and all-gather the
    dataset.  Returns the full dataset on every rank as torch tensors on ``device``."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = setpoints.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    res = run_local(setpoints[lo:hi], disturbances[lo:hi])
    local = {}
    for k in keys:
        v = res[k]
        local[k] = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(v), device=device)
    return gather_chunks(local, n, group=group, keys=keys)
