"""Data-parallel plumbing: one process per GPU, images sharded contiguously across ranks, no data-path
collective for inference.  torch.distributed is only used for the rendezvous, barriers and
the max-over-ranks reduction of device-timed intervals."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of n_items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def throughput(images_per_rank_step, steps, elapsed_ms_local, device="cpu"):
    """Whole-job images/s: all ranks' images over the slowest rank's device time."""
    total = sum_over_ranks(images_per_rank_step * steps, device)
    return total / (max_over_ranks(elapsed_ms_local, device) / 1e3)


def allreduce_mean_grads(grads, params):
    """exactly ONE all-reduce (sum, then / world) per training step over the flat gradient buffer.
    grads: {id(param): tensor}; params: parameters in a fixed (rank-independent) order.  Returns the averaged flat
    buffer and a {id(param): view} dict into it (identity copy when not initialised / world 1)."""
    order = [p for p in params if id(p) in grads]
    flat = torch.cat([grads[id(p)].reshape(-1).float() for p in order])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
        flat.div_(dist.get_world_size())
    views, off = {}, 0
    for p in order:
        n = grads[id(p)].numel()
        views[id(p)] = flat[off:off + n].view(grads[id(p)].shape)
        off += n
    return flat, views
