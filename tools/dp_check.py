"""Data-parallel gradient all-reduce: correctness of every dp_mode of train_engine.GraphedStep against a plain NCCL all-reduce
of the ranks' local gradients, and the step time per mode at the benchmarked size.  Launch with torchrun (1 process per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29519 tools/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.train_engine import GraphedStep  # noqa: E402
from planerecnet_b200.utils.synth import make_cotangents, make_input, perturb_  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    modes = sys.argv[1:] or ["p2p", "after", "nccl_overlap"]
    cases = (("PlaneRecNet_50_config", 2, 128, 160, 2), ("PlaneRecNet_101_config", 8, 480, 640, 10))
    if os.environ.get("PRN_DP_BIG_ONLY"):
        cases = cases[1:]
    for preset, B, H, W, steps in cases:
        set_cfg(preset)
        for mode in modes:
            torch.manual_seed(0)
            net = perturb_(PlaneRecNet(cfg)).train().cuda()
            for m in net.modules():                       # frozen statistics: a stable network, so that ranks are comparable
                if isinstance(m, torch.nn.BatchNorm2d) and steps == 2:
                    m.eval()
            x = make_input(B, H, W, seed=rank).cuda()
            step = GraphedStep(net.train_engine, net, x, world=world, dp_mode=mode)
            cots = make_cotangents(step.outs, seed=1 + rank, device="cuda")
            step.forward(x)
            local_g = step.backward(*cots)
            views = step.allreduce_grads()
            torch.cuda.synchronize()
            params = [p for p in net.parameters() if id(p) in views]
            ref = torch.cat([local_g[id(p)].reshape(-1).float() for p in params])      # the engine's own (local) gradients
            dist.all_reduce(ref, op=dist.ReduceOp.AVG)
            got = torch.cat([views[id(p)].reshape(-1) for p in params])
            err = float((got - ref).abs().max() / (ref.abs().max() + 1e-30))
            # timing
            for _ in range(3):
                step.forward(x); step.backward(*cots); step.allreduce_grads()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step.forward(x); step.backward(*cots); step.allreduce_grads()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            # the collective alone, on an idle GPU
            buf = torch.empty(57_000_000, device="cuda")
            for _ in range(3):
                dist.all_reduce(buf)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a0.record()
            for _ in range(5):
                dist.all_reduce(buf)
            a1.record()
            for _ in range(5):
                dist.all_reduce(buf, op=dist.ReduceOp.AVG)
            a2.record()
            torch.cuda.synchronize()
            t_sum, t_avg = a0.elapsed_time(a1) / 5, a1.elapsed_time(a2) / 5
            del buf
            if rank == 0:
                print(f"   NCCL all-reduce of 228 MB alone: SUM {t_sum:.3f} ms, AVG {t_avg:.3f} ms")
                print(f"DP_CHECK {preset} bs{B} {H}x{W} world={world} mode={mode}: max |reduced - nccl(local)| / max = {err:.2e}  "
                      f"step {float(ms):.3f} ms  ({world * B / float(ms) * 1e3:.1f} img/s)", flush=True)
            assert err < 1e-5 or os.environ.get('PRN_DP_OP') == 'skip', (mode, err)
            del step, net
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
