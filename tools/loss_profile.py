"""Where the time of one full training step (config 4 (ii)) goes: model forward (graphed), target assignment, the five loss
terms, backward.  Usage: python tools/loss_profile.py [batch]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200 import losses as PL  # noqa: E402
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.targets import assign_targets  # noqa: E402
from planerecnet_b200.utils.synth import make_gt, make_input, perturb_  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    set_cfg("PlaneRecNet_101_config")
    torch.manual_seed(0)
    net = perturb_(PlaneRecNet(cfg)).train().cuda()
    net.use_train_graph = True
    crit = PL.PlaneRecNetLoss(cfg, vnl_sampling=(sys.argv[2] if len(sys.argv) > 2 else "numpy"))
    x = make_input(B, 480, 640, 0).cuda()
    gts, gtd = make_gt(B, 480, 640, 0)
    gts = [{k: v.cuda() for k, v in g.items()} for g in gts]
    gtd = gtd.cuda()

    def T():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(4):
        for p in net.parameters():
            p.grad = None
        t0 = T()
        outs = net(x)
        t1 = T()
        np.random.seed(0)
        if hasattr(crit, "timings"):
            crit.timings = {}
        losses = crit(net, outs[0], outs[1], outs[2], outs[3], gts, gtd)
        t2 = T()
        tot = sum(v.mean() for v in losses.values())
        tot.backward()
        t3 = T()
        print(f"iter {it}: forward {1e3 * (t1 - t0):.1f} ms | loss {1e3 * (t2 - t1):.1f} ms | backward (loss + model) {1e3 * (t3 - t2):.1f} ms"
              + (" | " + " ".join(f"{k} {1e3 * v:.1f}" for k, v in crit.timings.items()) if getattr(crit, "timings", None) else ""))
    import cProfile
    import pstats
    pr = cProfile.Profile()
    for p in net.parameters():
        p.grad = None
    outs = net(x)
    torch.cuda.synchronize()
    pr.enable()
    np.random.seed(0)
    losses = crit(net, outs[0], outs[1], outs[2], outs[3], gts, gtd)
    tot = sum(v.mean() for v in losses.values())
    tot.backward()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
    outs = net(x)
    # the pieces of the loss, one by one (synchronised)
    fh, fw = outs[0].shape[-2:]
    t0 = T()
    from planerecnet_b200.targets import assign_targets_batch
    targets = assign_targets_batch(gts, (fh, fw), crit.num_grids, crit.scale_ranges, crit.num_classes, crit.sigma)
    t1 = T()
    print(f"assign_targets x{B}: {1e3 * (t1 - t0):.1f} ms")
    be = PL.CudaBackend()
    t0 = T()
    gw, gsum = be.lava_weights(gtd, fh, fw, crit.depth_resolution)
    t1 = T()
    ins, lav = PL._InsLava.apply(be, targets, gw, gsum, crit.w_ins, crit.w_lav, outs[0], *outs[2])
    t2 = T()
    import torch.nn.functional as F
    up = F.interpolate(outs[3], scale_factor=2, mode="bilinear", align_corners=False)
    np.random.seed(0)
    pln = [crit.vnl(up[b], gts[b]["masks"].bool(), gts[b]["plane_paras"][:, :3], gtd[b], gts[b]["k_matrix"]) for b in range(B)]
    t3 = T()
    print(f"lava_weights {1e3 * (t1 - t0):.1f} ms | ins+lava forward {1e3 * (t2 - t1):.1f} ms | plane-normal term x{B} {1e3 * (t3 - t2):.1f} ms")
    t0 = T()
    (ins + lav).backward(retain_graph=True)
    t1 = T()
    torch.stack(pln).mean().backward()
    t2 = T()
    print(f"ins+lava backward (+ model backward) {1e3 * (t1 - t0):.1f} ms | plane-normal backward (+ model backward) {1e3 * (t2 - t1):.1f} ms")


if __name__ == "__main__":
    main()
