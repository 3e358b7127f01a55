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
    for preset, B, H, W, steps in (("PlaneRecNet_50_config", 2, 128, 160, 2), ("PlaneRecNet_101_config", 8, 480, 640, 10)):
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
            if rank == 0:
                print(f"DP_CHECK {preset} bs{B} {H}x{W} world={world} mode={mode}: max |reduced - nccl(local)| / max = {err:.2e}  "
                      f"step {float(ms):.3f} ms  ({world * B / float(ms) * 1e3:.1f} img/s)", flush=True)
            assert err < 1e-5, (mode, err)
            del step, net
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
