"""Utterance sharding across GPUs: one process per GPU, contiguous block shards, NO collective inside the sampler
(every utterance's trajectory is independent — SURVEY.md §8e).  torch.distributed is used only for rendezvous,
the weight broadcast, the timing barrier and the max-over-ranks reduction of the measured time."""
import os

import torch
import torch.distributed as dist


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(rank, world, total):
    """Contiguous block [lo, hi) of `total` utterances owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def gather_utterances(local, total, device="cpu"):
    """Concatenate every rank's [n_local, N] result on all ranks, in global utterance order (outside the timed loop)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(r, world, total) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    mine = torch.zeros(nmax, local.shape[1], device=device, dtype=local.dtype)
    mine[:local.shape[0]] = local.to(device)
    bufs = [torch.empty_like(mine) for _ in sizes]       # all_gather needs equal sizes: pad, then trim
    dist.all_gather(bufs, mine)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
