"""world_size-2 gloo test of the data-parallel plumbing used by bench.py --gpus N (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from planerecnet_b200.utils import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert D.env_rank_world() == (rank, world, rank)
    lo, hi = D.shard_range(17, rank, world)
    D.barrier()
    ms = D.max_over_ranks(10.0 + 5.0 * rank)          # slowest rank defines the step time
    thr = D.throughput(8, 4, 100.0 * (rank + 1))       # 2 ranks x 8 images x 4 steps over 200 ms
    # the training step's single gradient all-reduce (mean over ranks) on a toy parameter set
    params = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
    grads = {id(params[0]): torch.full((3, 2), float(rank + 1)), id(params[1]): torch.arange(5.0) * (rank + 1)}   # params[2]: no grad
    flat, views = D.allreduce_mean_grads(grads, params)
    ok = (flat.numel() == 11 and torch.allclose(views[id(params[0])], torch.full((3, 2), 1.5)) and
          torch.allclose(views[id(params[1])], torch.arange(5.0) * 1.5) and id(params[2]) not in views)
    out.put((rank, lo, hi, ms, thr, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[5] for r in res)
    (_, lo0, hi0, ms0, thr0, _), (_, lo1, hi1, ms1, thr1, _) = res
    assert (lo0, hi0, lo1, hi1) == (0, 9, 9, 17)          # contiguous, covers everything, sizes differ by <= 1
    assert ms0 == ms1 == 15.0
    assert abs(thr0 - 320.0) < 1e-9 and thr0 == thr1      # 64 images / 0.2 s


def test_single_process_identities():
    assert D.shard_range(8, 0, 1) == (0, 8)
    assert D.max_over_ranks(3.5) == 3.5 and D.sum_over_ranks(2.0) == 2.0
    assert [D.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_peer_chunks_partition_every_slice():
    """utils.dist.peer_chunks (the reduce-scatter layout of PeerAllReduce): consecutive, disjoint, covering, 16-byte aligned starts."""
    from planerecnet_b200.utils.dist import peer_chunks
    for lo, hi in ((0, 14_000_003), (4096, 4096 + 7), (100, 100), (8, 1_500_000)):
        for world in (2, 4, 8):
            ch = peer_chunks(lo, hi, world)
            assert len(ch) == world and ch[0][0] == lo and ch[-1][1] == hi
            for (a0, a1), (b0, b1) in zip(ch, ch[1:]):
                assert a1 == b0 and a0 <= a1
            assert all((a0 - lo) % 4 == 0 for a0, a1 in ch if a1 > a0)
            assert sum(a1 - a0 for a0, a1 in ch) == hi - lo
