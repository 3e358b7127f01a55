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


def peer_chunks(lo, hi, world):
    """[lo, hi) cut into `world` consecutive chunks (16-byte aligned lengths; trailing chunks may be empty): chunk r is the part
    rank r reduces in PeerAllReduce."""
    n = hi - lo
    cs = (n + world - 1) // world
    cs = (cs + 3) // 4 * 4
    return [(min(lo + r * cs, hi), min(lo + (r + 1) * cs, hi)) for r in range(world)]


class PeerAllReduce:
    """Mean all-reduce of slices of a persistent fp32 buffer over the ranks of ONE node through peer memory (NVLink / NVSwitch)
    with the copy engines doing the transfers — no communication kernel occupies an SM for the duration of the collective.

    Why not NCCL here: the conv kernels are persistent, one CTA per SM with the whole register file; an NCCL all-reduce that is
    overlapped with the backward gets its CTAs onto SMs between two kernels and then keeps every following 148-CTA launch from
    completing until the collective is over (measured at N = 2: overlapped NCCL buckets made the step 4.6 % slower, not faster).
    Copies through `cudaMemcpyPeerAsync` need no SM; what is left on the SMs is a microsecond-scale sum kernel per bucket and
    4-byte NCCL all-reduces used as device-side barriers between the ranks' communication streams.

    Per slice [lo, hi), with chunk r = the r-th of `world` equal parts:
      barrier (every rank's gradients of this bucket are final)  ->  pull chunk `rank` of every peer's buffer into staging rows
      ->  own chunk = (own + sum of the rows) / world  ->  barrier (reduced chunks are final)  ->  pull every peer's reduced chunk.
    `finish()` is one more barrier: nobody still reads this rank's buffer when the next step overwrites it.
    All calls enqueue on the CURRENT stream (the caller's communication stream)."""

    def __init__(self, flat, slices, group=None):
        import torch.multiprocessing.reductions as R
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        self.flat, self.slices = flat, list(slices)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group
        fn, args = R.reduce_tensor(flat)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (fn, args), group=group)
        self.peers = [flat if r == self.rank else gathered[r][0](*gathered[r][1]) for r in range(self.world)]
        for r, t in enumerate(self.peers):
            assert t.numel() == flat.numel() and t.dtype == flat.dtype, f"rank {r} exposes a different gradient buffer"
        self.chunk = [self._chunks(lo, hi) for lo, hi in self.slices]
        longest = max((b - a for ch in self.chunk for a, b in ch), default=0)
        self.staging = torch.empty(max(self.world - 1, 1), max(longest, 1), dtype=torch.float32, device=flat.device)
        self.token = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self.order = [(self.rank + k) % self.world for k in range(1, self.world)]      # spread the pulls over the peers

    def _chunks(self, lo, hi):
        return peer_chunks(lo, hi, self.world)

    def barrier(self):
        dist.all_reduce(self.token, group=self.group)                  # 4 bytes: a device-side barrier between the ranks' streams

    def reduce(self, b):
        ch = self.chunk[b]
        a0, a1 = ch[self.rank]
        n = a1 - a0
        self.barrier()
        if n > 0:
            for j, p in enumerate(self.order):
                self.staging[j, :n].copy_(self.peers[p][a0:a1], non_blocking=True)
            mine = self.flat[a0:a1]
            mine.add_(self.staging[:self.world - 1, :n].sum(0)).div_(self.world)
        self.barrier()
        for p in self.order:
            c0, c1 = ch[p]
            if c1 > c0:
                self.flat[c0:c1].copy_(self.peers[p][c0:c1], non_blocking=True)

    def finish(self):
        self.barrier()
